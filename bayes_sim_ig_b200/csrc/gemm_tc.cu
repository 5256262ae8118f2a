// tcgen05 tensor-core GEMM engine (BSIG_GEMM_TC_TF32 / BSIG_GEMM_TC_TF32X3) for
// the dense contractions of the path: the RFF projection (models/rff.py:128-132)
// and the MLP / head layers at large batch (models/mdnn.py:108-119).
//
//   C[M,N] = epi( sum_r A(i,r) * B(r,j) ),  every operand row-major in EITHER orientation:
//   K-major (the reduction index contiguous: x . W^T forward layers, the RFF projection) or
//   MN-major (the output index contiguous: dgrad reads W[n][k] along k, wgrad reads
//   dY[b][n] / X[b][k] along n / k) -- so forward, dgrad and wgrad all run on tcgen05
//   straight from the natural row-major tensors (UMMA descriptors carry the major-ness).
//
// Blackwell-native structure: TMA (cp.async.bulk.tensor, 128B swizzle) stages
// 128x32 fp32 operand tiles in shared memory, ONE elected thread issues
// tcgen05.mma.kind::tf32 (M=128, N=128, K=8) accumulating fp32 in tensor memory,
// completion is tracked with mbarriers (tcgen05.commit), the epilogue warps read
// the accumulator with tcgen05.ld and apply bias / tanh / scaled cos|sin.
//
// fp32 parity (TF32X3): the fp32 operands are split in shared memory into
// hi = tf32(a) and lo = a - hi by the converter warps, and every K step issues
// three MMAs (hi*hi + hi*lo + lo*hi): the dropped lo*lo term is 2^-22 relative,
// so the result matches an fp32 FFMA GEMM to ~1e-6.  Plain TF32 issues one MMA.
//
// TMA needs 16-byte aligned bases and row pitches and cannot gather rows: operands that
// do not qualify (feature widths = 2 mod 4 such as 302 / 11 802 / 105 002, the 270-wide head
// output, minibatch row gathers) are first copied into a padded staging buffer in the
// caller's workspace (stage_rows_kernel).  Skinny outputs with a long reduction (weight
// gradients at large batch, the ShadowHand first layer) split K over the CTAs; the fp32
// partials are reduced in fixed order by splitk_reduce_kernel (gemm_simt.cu), which also
// applies the epilogue.
#include <cuda.h>

#include <algorithm>

#include "common.cuh"
#include "gemm.cuh"
#include "tc_ptx.cuh"

namespace bsig {

namespace tc {

constexpr int BM = 128, BN = 128, BK = 32;       // BK fp32 = one 128-byte swizzle row
constexpr int STAGES_X1 = 6, STAGES_X3 = 3;       // 192 KB of operand stages per (persistent) CTA
constexpr int TILE_BYTES = BM * BK * 4;           // 16 KB
constexpr int NUM_THREADS = 320;                  // warp0 TMA, warp1 MMA, warps2-9 convert+epilogue

struct TcArgs {
  float* C;
  int64_t ldc;
  const float* bias;
  const float* aux;        // EPI_MUL_DTANH: h of the previous layer [M, ld_aux]
  int64_t ld_aux;
  int M, N, K;
  int epi;
  float scale;
  int a_mn, b_mn;          // operand orientation: 0 = K-major, 1 = MN-major
  int splits;              // K split over this many work items per output tile
  int kb_per_split;        // BK-blocks per split
  float* partial;          // splits > 1: raw accumulators [splits][M][N]
};

// Persistent, warp-specialised kernel: one CTA per SM walks the output tiles.
//   warp 0      TMA producer          (smem ring of STAGES stages)
//   warp 1      MMA issuer            (two 128-column TMEM accumulators, ping-pong)
//   warps 2-5   X3: hi/lo converters; X1: second epilogue group
//   warps 6-9   epilogue              (tcgen05.ld -> bias/tanh/sincos -> global)
// The epilogue of tile t overlaps the TMA + MMA mainloop of tile t+1.
template <bool X3>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               TcArgs g) {
  constexpr int MAX_STAGES = X3 ? STAGES_X3 + 1 : STAGES_X1;
  constexpr int EPI_WARPS = X3 ? 4 : 8;
  constexpr uint32_t TMEM_COLS = X3 ? 512 : 256;
  // X3 with a K-major A: the converter warps write hi / lo of A into tensor memory and the
  // MMA takes A from there -- the tensor core then reads only B from shared memory, whose
  // bandwidth (128 B/clk: exactly what three SS MMAs per K step consume) was the limit of
  // the all-shared-memory form
  const bool a_ts = X3 && !g.a_mn;
  // stage = A, B (+ A_lo, B_lo); with A in tensor memory the A_lo tile does not exist and the
  // same 192 KB hold FOUR stages (the TF32x3 form is bound by the TMA -> convert -> MMA ->
  // release latency chain, i.e. by the ring depth)
  const int TILES_PER_STAGE = X3 ? (a_ts ? 3 : 4) : 2;
  const int STAGES = X3 ? (a_ts ? STAGES_X3 + 1 : STAGES_X3) : STAGES_X1;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // (offset arithmetic on the array keeps the shared address space: LDS / STS, not generic)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t full_bar[MAX_STAGES], conv_bar[MAX_STAGES], empty_bar[MAX_STAGES];
  __shared__ uint64_t tmem_full_bar[2], tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_kb = (g.K + BK - 1) / BK;
  const int n_tiles = (g.N + BN - 1) / BN;
  const int num_tiles = ((g.M + BM - 1) / BM) * n_tiles;
  const int num_items = num_tiles * g.splits;      // work item = (output tile, K split)
  constexpr uint32_t MN_BOX = (uint32_t)BK * 128u;   // bytes of one [32 x BK] MN-major box

  auto tile_a = [&](int s) { return smem + (size_t)s * TILES_PER_STAGE * TILE_BYTES; };
  auto tile_b = [&](int s) { return tile_a(s) + TILE_BYTES; };
  auto tile_alo = [&](int s) { return tile_a(s) + 2 * TILE_BYTES; };
  auto tile_blo = [&](int s) { return tile_a(s) + (TILES_PER_STAGE - 1) * TILE_BYTES; };

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&conv_bar[s], 128);
      mbar_init(&empty_bar[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], EPI_WARPS * 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    // two ping-pong accumulators of 128 fp32 columns; X3 additionally stages the hi / lo
    // halves of the A operand of every pipeline stage in tensor memory (64 columns each)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(&tmem_base_slot)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;                         // ring position and phase (no division per stage)
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        const int tile = item / g.splits, split = item - tile * g.splits;
        const int m0 = (tile / n_tiles) * BM, n0 = (tile % n_tiles) * BN;
        const int kb0 = split * g.kb_per_split, kb1 = min(num_kb, kb0 + g.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb, ph ^= (uint32_t)(++s == STAGES), s = (s == STAGES ? 0 : s)) {
          mbar_wait(&empty_bar[s], ph ^ 1);
          mbar_expect_tx(&full_bar[s], 2 * TILE_BYTES);
          if (!g.a_mn) {
            tma_load_2d(tile_a(s), &map_a, &full_bar[s], kb * BK, m0);
          } else {
#pragma unroll
            for (int blk = 0; blk < BM / 32; ++blk)
              tma_load_2d(tile_a(s) + blk * MN_BOX, &map_a, &full_bar[s], m0 + 32 * blk, kb * BK);
          }
          if (!g.b_mn) {
            tma_load_2d(tile_b(s), &map_b, &full_bar[s], kb * BK, n0);
          } else {
#pragma unroll
            for (int blk = 0; blk < BN / 32; ++blk)
              tma_load_2d(tile_b(s) + blk * MN_BOX, &map_b, &full_bar[s], n0 + 32 * blk, kb * BK);
          }
        }
      }
    }
  } else if (warp == 1) {
    // -------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      // instruction descriptor: D=f32, A=B=tf32, K-major both, N=128, M=128
      // (bits 15 / 16: A / B are MN-major)
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(g.a_mn & 1) << 15) |
                             ((uint32_t)(g.b_mn & 1) << 16) | ((uint32_t)(BN >> 3) << 17) |
                             ((uint32_t)(BM >> 4) << 24);
      int s = 0, tcount = 0;
      uint32_t ph = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++tcount) {
        const int split = item % g.splits;
        const int kb0 = split * g.kb_per_split, kb1 = min(num_kb, kb0 + g.kb_per_split);
        const int buf = tcount & 1;
        mbar_wait(&tmem_empty_bar[buf], ((tcount >> 1) & 1) ^ 1);   // epilogue drained it
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d_tmem = tmem_base + (uint32_t)(buf * BN);
        for (int kb = kb0; kb < kb1; ++kb, ph ^= (uint32_t)(++s == STAGES), s = (s == STAGES ? 0 : s)) {
          const uint32_t parity = ph;
          if (X3) mbar_wait(&conv_bar[s], parity);
          else mbar_wait(&full_bar[s], parity);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          // The issuing thread is ONE chain of dependent instructions (~5 cycles each): the four
          // operand descriptors of a stage are built once, a K step only adds its offset to the
          // address field (32 bytes along a K-major swizzled row = +2 descriptor units, one
          // 1024-byte atom of an MN-major tile = +64; tile bases are 1024-byte aligned, so the
          // 14-bit field never carries)
          const uint32_t a_hi = smem_u32(tile_a(s)), b_hi = smem_u32(tile_b(s));
          const uint64_t a_step = g.a_mn ? 64u : 2u, b_step = g.b_mn ? 64u : 2u;
          const uint64_t db_hi = g.b_mn ? umma_desc_mn(b_hi, MN_BOX) : umma_desc(b_hi);
          if (a_ts) {
            const uint32_t b_lo = smem_u32(tile_blo(s));
            const uint64_t db_lo = umma_desc(b_lo);            // (a_ts implies X3; B may be MN-major)
            const uint64_t db_lo2 = g.b_mn ? umma_desc_mn(b_lo, MN_BOX) : db_lo;
            const uint32_t a_tm0 = tmem_base + 256u + (uint32_t)(s * 64);
#pragma unroll
            for (int k = 0; k < BK / 8; ++k) {
              const uint32_t acc = (kb > kb0 || k > 0) ? 1u : 0u;
              const uint32_t a_tm = a_tm0 + (uint32_t)(k * 8);
              umma_tf32_ts(d_tmem, a_tm, db_hi + k * b_step, idesc, acc);
              umma_tf32_ts(d_tmem, a_tm, db_lo2 + k * b_step, idesc, 1u);
              umma_tf32_ts(d_tmem, a_tm + 32u, db_hi + k * b_step, idesc, 1u);
            }
          } else {
            const uint64_t da_hi = g.a_mn ? umma_desc_mn(a_hi, MN_BOX) : umma_desc(a_hi);
            uint64_t da_lo = 0, db_lo = 0;
            if (X3) {
              const uint32_t a_lo = smem_u32(tile_alo(s)), b_lo = smem_u32(tile_blo(s));
              da_lo = g.a_mn ? umma_desc_mn(a_lo, MN_BOX) : umma_desc(a_lo);
              db_lo = g.b_mn ? umma_desc_mn(b_lo, MN_BOX) : umma_desc(b_lo);
            }
#pragma unroll
            for (int k = 0; k < BK / 8; ++k) {
              const uint32_t acc = (kb > kb0 || k > 0) ? 1u : 0u;
              umma_tf32(d_tmem, da_hi + k * a_step, db_hi + k * b_step, idesc, acc);
              if (X3) {
                umma_tf32(d_tmem, da_hi + k * a_step, db_lo + k * b_step, idesc, 1u);
                umma_tf32(d_tmem, da_lo + k * a_step, db_hi + k * b_step, idesc, 1u);
              }
            }
          }
          umma_commit(&empty_bar[s]);           // stage reusable once these MMAs retire
        }
        umma_commit(&tmem_full_bar[buf]);       // accumulator of this tile complete
      }
    }
  } else if (X3 && warp < 6) {
    // ------------------------------------------------- hi/lo converters (X3 only)
    const int t = threadIdx.x - 64;             // 0..127
    int s = 0;
    uint32_t ph = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      const int split = item % g.splits;
      const int kb0 = split * g.kb_per_split, kb1 = min(num_kb, kb0 + g.kb_per_split);
      for (int kb = kb0; kb < kb1; ++kb, ph ^= (uint32_t)(++s == STAGES), s = (s == STAGES ? 0 : s)) {
        mbar_wait(&full_bar[s], ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        float4* a = reinterpret_cast<float4*>(tile_a(s));
        float4* b = reinterpret_cast<float4*>(tile_b(s));
        float4* alo = reinterpret_cast<float4*>(tile_alo(s));
        float4* blo = reinterpret_cast<float4*>(tile_blo(s));
        auto split = [](float4* hi_p, float4* lo_p, int idx) {
          const float4 v = hi_p[idx];
          float4 h, l;
          h.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u);
          h.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
          h.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u);
          h.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
          l.x = v.x - h.x; l.y = v.y - h.y; l.z = v.z - h.z; l.w = v.w - h.w;
          hi_p[idx] = h;
          lo_p[idx] = l;
        };
        if (a_ts) {
          // thread <-> row of the A tile (the TMEM lane quadrant of a warp is warp % 4):
          // read the row's 32 K values (128 B, 16-byte chunks XOR-swizzled with row % 8),
          // split, and store hi / lo as 32 + 32 tensor-memory columns of this stage
          const int m = (warp & 3) * 32 + lane;
          const uint8_t* arow = tile_a(s) + m * 128;
          uint32_t hi[32], lo[32];
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float4 v = *reinterpret_cast<const float4*>(arow + ((c ^ (m & 7)) << 4));
            const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const uint32_t h = __float_as_uint(vv[u]) & 0xffffe000u;
              hi[4 * c + u] = h;
              lo[4 * c + u] = __float_as_uint(vv[u] - __uint_as_float(h));
            }
          }
          const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + 256u + (uint32_t)(s * 64);
#define BSIG_ST32(ADDR, R)                                                                       \
  asm volatile(                                                                                  \
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "                                            \
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "                 \
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(  \
          ADDR),                                                                                 \
      "r"(R[0]), "r"(R[1]), "r"(R[2]), "r"(R[3]), "r"(R[4]), "r"(R[5]), "r"(R[6]), "r"(R[7]),    \
      "r"(R[8]), "r"(R[9]), "r"(R[10]), "r"(R[11]), "r"(R[12]), "r"(R[13]), "r"(R[14]),          \
      "r"(R[15]), "r"(R[16]), "r"(R[17]), "r"(R[18]), "r"(R[19]), "r"(R[20]), "r"(R[21]),        \
      "r"(R[22]), "r"(R[23]), "r"(R[24]), "r"(R[25]), "r"(R[26]), "r"(R[27]), "r"(R[28]),        \
      "r"(R[29]), "r"(R[30]), "r"(R[31])                                                         \
      : "memory")
          BSIG_ST32(taddr, hi);
          BSIG_ST32(taddr + 32u, lo);
#undef BSIG_ST32
          asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
#pragma unroll
          for (int q = 0; q < TILE_BYTES / 16 / 128; ++q) split(b, blo, t + 128 * q);
        } else {
#pragma unroll
          for (int q = 0; q < TILE_BYTES / 16 / 128; ++q) {   // 8 float4 per thread per tile
            split(a, alo, t + 128 * q);
            split(b, blo, t + 128 * q);
          }
        }
        // make the generic-proxy writes visible to the tensor core (async proxy), order the
        // tensor-memory stores before the MMA warp's reads
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        mbar_arrive(&conv_bar[s]);
      }
    }
  } else {
    // ---------------------------------------------------------------- epilogue
    // warp quadrant q = warp % 4 owns TMEM lanes [32q, 32q+32) = rows of the tile;
    // with two warps per quadrant (X1) they take alternate 32-column blocks
    const int q = warp & 3;
    const int half = X3 ? 0 : ((warp - 2) >> 2);
    constexpr int CB_STEP = X3 ? 1 : 2;
    const bool vec_ok = ((g.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.C) & 15) == 0) &&
                        (g.epi != EPI_SINCOS || (g.N & 3) == 0);
    int tcount = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++tcount) {
      const int tile = item / g.splits, split = item - tile * g.splits;
      const int m0 = (tile / n_tiles) * BM, n0 = (tile % n_tiles) * BN;
      const int buf = tcount & 1;
      mbar_wait(&tmem_full_bar[buf], (tcount >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int row = m0 + q * 32 + lane;
      const bool row_ok = row < g.M;
      float* crow = g.C + (int64_t)(row_ok ? row : 0) * g.ldc;
      const float* auxrow = g.aux + (int64_t)(row_ok ? row : 0) * g.ld_aux;
      float* prow = g.partial + ((int64_t)split * g.M + (row_ok ? row : 0)) * g.N;
#pragma unroll 1
      for (int cb = half; cb < BN / 32; cb += CB_STEP) {
        const int jbase = n0 + cb * 32;
        if (jbase >= g.N) break;                // warp-uniform
        uint32_t r[32];
        const uint32_t taddr =
            tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + cb * 32);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
              "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]),
              "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
              "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
              "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
              "=r"(r[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (!row_ok) continue;
        if (g.splits > 1) {
          // raw accumulators: the fixed-order reduction + epilogue run in a second kernel
#pragma unroll
          for (int c = 0; c < 32; ++c)
            if (jbase + c < g.N) prow[jbase + c] = __uint_as_float(r[c]);
          continue;
        }
        const bool full_block = jbase + 32 <= g.N;
        // four columns at a time, everything statically indexed (stays in registers)
#pragma unroll
        for (int c = 0; c < 32; c += 4) {
          float v[4], w2[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int j = min(jbase + c + u, g.N - 1);
            float acc = __uint_as_float(r[c + u]);
            w2[u] = 0.f;
            if (g.epi == EPI_BIAS || g.epi == EPI_BIAS_TANH) acc += __ldg(g.bias + j);
            if (g.epi == EPI_BIAS_TANH) acc = tanhf(acc);
            if (g.epi == EPI_MUL_DTANH) {
              const float h = __ldg(auxrow + j);
              acc *= (1.0f - h * h);
            }
            if (g.epi == EPI_SINCOS) {
              // two-term Cody-Waite reduction to [-pi, pi], then the SFU sin/cos
              // (abs error ~5e-7 on features of magnitude `scale`)
              const float kf = rintf(acc * 0.15915494309189535f);
              float red = fmaf(-kf, 6.2831854820251465f, acc);
              red = fmaf(-kf, -1.7484555e-7f, red);
              acc = g.scale * __cosf(red);
              w2[u] = g.scale * __sinf(red);
            }
            v[u] = acc;
          }
          if (vec_ok && (full_block || jbase + c + 4 <= g.N)) {
            *reinterpret_cast<float4*>(crow + jbase + c) = make_float4(v[0], v[1], v[2], v[3]);
            if (g.epi == EPI_SINCOS)
              *reinterpret_cast<float4*>(crow + g.N + jbase + c) =
                  make_float4(w2[0], w2[1], w2[2], w2[3]);
          } else {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              if (jbase + c + u < g.N) {
                crow[jbase + c + u] = v[u];
                if (g.epi == EPI_SINCOS) crow[g.N + jbase + c + u] = w2[u];
              }
            }
          }
        }
      }
      // this warp has read everything it needs from the accumulator
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(&tmem_empty_bar[buf]);
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS)
                 : "memory");
  }
}

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) ==
            cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// Row-major operand [outer][inner] with row pitch ld floats.  K-major: inner = K, box =
// BK x 128 rows; MN-major: inner = the M (or N) index, box = 32 x BK rows.  128B swizzle.
static int make_map(CUtensorMap* map, const float* base, int64_t outer, int64_t inner, int64_t ld,
                    bool mn_major) {
  EncodeTiledFn fn = encode_fn();
  BSIG_REQUIRE(fn != nullptr, "gemm_tc: cuTensorMapEncodeTiled is not available");
  const cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
  const cuuint32_t box[2] = {(cuuint32_t)(mn_major ? 32 : BK), (cuuint32_t)(mn_major ? BK : BM)};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims,
                        strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  BSIG_REQUIRE(r == CUDA_SUCCESS, "gemm_tc: cuTensorMapEncodeTiled failed (%d)", (int)r);
  return 0;
}

// dst[r][c] = src[(rows ? rows[r] : r) * src_ld + c] for c < width, zero up to dst_ld:
// the padded / gathered staging copy of an operand TMA cannot address directly.
__global__ void __launch_bounds__(256)
stage_rows_kernel(const float* __restrict__ src, int64_t src_ld, const int64_t* __restrict__ rows,
                  float* __restrict__ dst, int64_t dst_ld, int64_t n_rows, int64_t width) {
  const int64_t q4 = dst_ld >> 2;                    // dst_ld % 4 == 0
  const int64_t total = n_rows * q4;
  const bool vec = ((src_ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = e / q4, c = 4 * (e - r * q4);
    const float* sp = src + (rows ? __ldg(rows + r) : r) * src_ld + c;
    float4 v;
    if (vec && c + 4 <= width) {
      v = __ldg(reinterpret_cast<const float4*>(sp));
    } else {
      v.x = c + 0 < width ? __ldg(sp + 0) : 0.f;
      v.y = c + 1 < width ? __ldg(sp + 1) : 0.f;
      v.z = c + 2 < width ? __ldg(sp + 2) : 0.f;
      v.w = c + 3 < width ? __ldg(sp + 3) : 0.f;
    }
    *reinterpret_cast<float4*>(dst + r * dst_ld + c) = v;
  }
}

// How one operand reaches TMA: orientation, extent of the two row-major indices, pitch,
// gather of the outer index, and whether it must be staged first.
struct TcOperand {
  bool ok, mn_major, stage;
  const float* base;
  int64_t outer, inner, ld;
  const int64_t* gather;
};

static TcOperand classify(const float* base, int64_t mn_extent, int64_t K, int64_t s_mn, int64_t s_k,
                          const int64_t* gather_mn, const int64_t* gather_k) {
  TcOperand o = {};
  o.base = base;
  if (s_k == 1 && gather_k == nullptr) {             // K contiguous: rows = the M / N index
    o.ok = true; o.mn_major = false; o.outer = mn_extent; o.inner = K; o.ld = s_mn;
    o.gather = gather_mn;
  } else if (s_mn == 1 && gather_mn == nullptr) {    // M / N contiguous: rows = the K index
    o.ok = true; o.mn_major = true; o.outer = K; o.inner = mn_extent; o.ld = s_k;
    o.gather = gather_k;
  }
  if (o.ok)
    o.stage = o.gather != nullptr || (o.ld & 3) != 0 || o.ld < o.inner ||
              (reinterpret_cast<uintptr_t>(base) & 15) != 0;
  return o;
}

static int64_t staged_floats(int64_t outer, int64_t inner) { return outer * ((inner + 3) & ~(int64_t)3) + 64; }

// K split of a skinny problem: enough work items for one wave, at least four BK-blocks each
static int tc_splits(int64_t M, int64_t N, int64_t K) {
  const int64_t tiles = ceil_div(M, BM) * ceil_div(N, BN);
  const int64_t num_kb = ceil_div(K, BK);
  const int64_t sms = sm_count();
  if (tiles * 2 > sms || num_kb < 8) return 1;
  return (int)std::max<int64_t>(1, std::min<int64_t>(sms / tiles, num_kb / 4));
}

}  // namespace tc

bool gemm_tc_applicable(const GemmArgs& g) {
  const tc::TcOperand a = tc::classify(g.A, g.M, g.K, g.a_si, g.a_sr, g.a_rows, nullptr);
  const tc::TcOperand b = tc::classify(g.B, g.N, g.K, g.b_sj, g.b_sr, nullptr, g.b_rows);
  const bool epi_ok = g.epi == EPI_STORE || g.epi == EPI_BIAS || g.epi == EPI_BIAS_TANH ||
                      g.epi == EPI_MUL_DTANH || g.epi == EPI_SINCOS;
  return a.ok && b.ok && epi_ok && g.rowsum == nullptr && g.M >= 1 && g.N >= 1 && g.K >= 1;
}

// workspace: staging copies of both operands (worst case) + split-K partials
int64_t gemm_tc_ws_bytes(int64_t M, int64_t N, int64_t K) {
  const int64_t stage = std::max(tc::staged_floats(M, K), tc::staged_floats(K, M)) +
                        std::max(tc::staged_floats(N, K), tc::staged_floats(K, N));
  const int s = tc::tc_splits(M, N, K);
  return 4 * (stage + (s > 1 ? (int64_t)s * M * N : 0)) + 1024;
}

int gemm_tc(const GemmArgs& g, bool x3, void* ws, int64_t ws_bytes, cudaStream_t st) {
  using namespace tc;
  TcOperand a = classify(g.A, g.M, g.K, g.a_si, g.a_sr, g.a_rows, nullptr);
  TcOperand b = classify(g.B, g.N, g.K, g.b_sj, g.b_sr, nullptr, g.b_rows);
  BSIG_REQUIRE(a.ok && b.ok, "gemm_tc: operand layout not supported");
  float* wsp = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255);
  int64_t ws_left = ws == nullptr ? 0 : (ws_bytes - (reinterpret_cast<char*>(wsp) - (char*)ws)) / 4;
  auto stage = [&](TcOperand& o) -> int {
    if (!o.stage) return 0;
    const int64_t dst_ld = (o.inner + 3) & ~(int64_t)3;
    const int64_t need = staged_floats(o.outer, o.inner);
    BSIG_REQUIRE(ws_left >= need, "gemm_tc: workspace too small for the staging copy");
    const int64_t work = o.outer * (dst_ld >> 2);
    const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(work, 256), (int64_t)sm_count() * 8));
    stage_rows_kernel<<<blocks, 256, 0, st>>>(o.base, o.ld, o.gather, wsp, dst_ld, o.outer, o.inner);
    BSIG_LAUNCH_CHECK();
    o.base = wsp; o.ld = dst_ld; o.gather = nullptr; o.stage = false;
    wsp += need & ~(int64_t)63; ws_left -= need & ~(int64_t)63;
    return 0;
  };
  if (stage(a)) return 1;
  if (stage(b)) return 1;
  CUtensorMap map_a, map_b;
  if (make_map(&map_a, a.base, a.outer, a.inner, a.ld, a.mn_major)) return 1;
  if (make_map(&map_b, b.base, b.outer, b.inner, b.ld, b.mn_major)) return 1;
  TcArgs t;
  t.C = g.C; t.ldc = g.ldc; t.bias = g.bias; t.aux = g.aux; t.ld_aux = g.ld_aux;
  t.M = g.M; t.N = g.N; t.K = g.K; t.epi = g.epi; t.scale = g.scale;
  t.a_mn = a.mn_major ? 1 : 0; t.b_mn = b.mn_major ? 1 : 0;
  int splits = tc_splits(g.M, g.N, g.K);
  if (splits > 1 && ws_left < (int64_t)splits * g.M * g.N) splits = 1;
  const int64_t num_kb = ceil_div(g.K, BK);
  t.kb_per_split = (int)ceil_div(num_kb, splits);
  splits = (int)ceil_div(num_kb, t.kb_per_split);
  t.splits = splits;
  t.partial = splits > 1 ? wsp : nullptr;
  const int64_t num_items = ceil_div(g.M, BM) * ceil_div(g.N, BN) * splits;
  const dim3 grid((unsigned)std::min<int64_t>(num_items, sm_count()));
  const size_t smem = (size_t)(x3 ? STAGES_X3 * 4 : STAGES_X1 * 2) * TILE_BYTES + 1024;   // 192 KB either way
  if (x3) {
    BSIG_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)smem));
    gemm_tc_kernel<true><<<grid, NUM_THREADS, smem, st>>>(map_a, map_b, t);
  } else {
    BSIG_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<false>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gemm_tc_kernel<false><<<grid, NUM_THREADS, smem, st>>>(map_a, map_b, t);
  }
  BSIG_LAUNCH_CHECK();
  if (splits > 1) {
    GemmArgs r = g;
    r.partial = t.partial;
    return gemm_splitk_reduce(r, splits, st);
  }
  return 0;
}

}  // namespace bsig
