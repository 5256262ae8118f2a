// Fused cross-correlation summarizer -> first dense layer.
//
// The reference materialises  x[n, p*Q + q] = sf[n,p] * af[n,q]  (+ mean, std of sf) per
// trajectory (utils/summarizers.py:106-119: 420 KB per ShadowHand trajectory) and feeds it to
// the first nn.Linear of the MDNN (models/mdnn.py:108).  x is a rank-one outer product, so
// nothing but its two factors ever has to exist in memory:
//
//   corr_factors_kernel   rollouts -> fac[n] = [ sf (S) | af (Q) | mean | std ]   (~4.6 KB)
//   corr_fwd_kernel       y = act(x W^T + b): split-K over the SMs; every CTA GENERATES its
//                         128 x 32 tiles of x (as tf32 hi / lo halves, straight into tensor
//                         memory) from the factors in shared memory, W tiles arrive by 8-byte
//                         cp.async into the swizzled operand layout (rows are 8-byte aligned),
//                         tcgen05.mma (TS form, TF32x3) accumulates in tensor memory
//   corr_wgrad_kernel     dW[j, k] = sum_n dy[n, j] x[n, k]: dy^T lives in tensor memory for the
//                         whole launch, the x tiles are generated into swizzled shared memory,
//                         and the epilogue applies Adam to W / exp_avg / exp_avg_sq in place
//                         (or stores dW for the data-parallel exchange) -- the 13.4 M-parameter
//                         ShadowHand weight gradient never touches HBM
//
// x values are the same fp32 products the summarizer kernel would have stored (one rounding),
// the TF32x3 split recovers the fp32 mantissas of both operands, accumulation is fp32 in TMEM.
#include <cuda.h>

#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "gemm.cuh"
#include "tc_ptx.cuh"

namespace bsig {
namespace corr {

using namespace tc;

// exact n / d for n, d < 2^20 (m = ceil(2^40 / d))
struct FastDiv {
  uint32_t d;
  uint64_t m;
};
__device__ __forceinline__ uint32_t fdiv(uint32_t n, const FastDiv& f) {
  return (uint32_t)(((uint64_t)n * f.m) >> 40);
}
static FastDiv mk_div(uint64_t d) {
  FastDiv f;
  f.d = (uint32_t)d;
  f.m = ((1ull << 40) + d - 1) / d;
  return f;
}

// ------------------------------------------------------------------ factors
struct FactorArgs {
  const float* states;
  const float* actions;
  float* fac;
  int64_t ldf;
  int* flag;
  int64_t n;
  int64_t s_stride, a_stride, s_tstride, a_tstride;
  int D, A, Pn, Qn, use_diff;
};

// One warp per trajectory.  mean / unbiased std of sf are formed exactly as crosscorr_kernel
// (summarizers.cu) forms them -- fp64, two passes, sequential for Pn <= 64 and lane-strided +
// butterfly otherwise -- so the two statistics are bit-identical to the materialised summary.
__global__ void __launch_bounds__(256) corr_factors_kernel(FactorArgs p, FastDiv divD, FastDiv divA) {
  extern __shared__ float fsm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t traj = (int64_t)blockIdx.x * 8 + warp;
  if (traj >= p.n) return;
  float* v = fsm + (size_t)warp * p.Pn;
  float* out = p.fac + traj * p.ldf;
  const int Dm1 = p.D - 1;
  bool bad = false;
  float smax = 0.f, amax = 0.f;
  for (int i = lane; i < p.Pn; i += 32) {
    const uint32_t t = fdiv((uint32_t)i, divD), j = (uint32_t)i - t * (uint32_t)Dm1;
    const float* s = p.states + traj * p.s_stride + t * p.s_tstride + j;
    const float lo = __ldg(s);
    const float x = p.use_diff ? (__ldg(s + 1) - lo) : lo;
    bad |= !finite_f(x);
    smax = fmaxf(smax, fabsf(x));
    v[i] = x;
    out[i] = x;
  }
  for (int q = lane; q < p.Qn; q += 32) {
    const uint32_t tq = fdiv((uint32_t)q, divA);
    const float x = __ldg(p.actions + traj * p.a_stride + tq * p.a_tstride +
                          ((uint32_t)q - tq * (uint32_t)p.A));
    bad |= !finite_f(x);
    amax = fmaxf(amax, fabsf(x));
    out[p.Pn + q] = x;
  }
  __syncwarp();
  float mean_f, sd_f;
  if (p.Pn <= 64) {
    double acc = 0.0;
    for (int i = 0; i < p.Pn; ++i) acc += (double)v[i];
    const double mean = acc / (double)p.Pn;
    double sq = 0.0;
    for (int i = 0; i < p.Pn; ++i) {
      const double dlt = (double)v[i] - mean;
      sq += dlt * dlt;
    }
    mean_f = (float)mean;
    sd_f = p.Pn < 2 ? 0.f : (float)sqrt(sq / (double)(p.Pn - 1));
  } else {
    double acc = 0.0;
    for (int i = lane; i < p.Pn; i += 32) acc += (double)v[i];
    acc = warp_sum(acc);
    const double mean = acc / (double)p.Pn;
    double sq = 0.0;
    for (int i = lane; i < p.Pn; i += 32) {
      const double dlt = (double)v[i] - mean;
      sq += dlt * dlt;
    }
    sq = warp_sum(sq);
    mean_f = (float)mean;
    sd_f = (float)sqrt(sq / (double)(p.Pn - 1));
  }
  // torch.isfinite(feats).all(): with finite factors a product can only overflow, and the
  // largest product of a row is max|sf| * max|af| (rounding is monotonic)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    smax = fmaxf(smax, __shfl_xor_sync(0xffffffffu, smax, o));
    amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
  }
  bad |= !finite_f(sd_f) || !finite_f(mean_f) || !(smax * amax <= 3.402823466e38f);
  if (lane == 0) {
    out[p.Pn + p.Qn] = mean_f;
    out[p.Pn + p.Qn + 1] = sd_f;
    for (int64_t c = p.Pn + p.Qn + 2; c < p.ldf; ++c) out[c] = 0.f;
  }
  if (__any_sync(0xffffffffu, bad) && lane == 0) atomicOr(p.flag, 1);
}

// ------------------------------------------------------------------ forward
constexpr int BM = 128, BN = 128, BK = 32;
constexpr int TILE_BYTES = BM * BK * 4;           // 16 KB
// warps 0-3 stream W (cp.async), warp 4 issues the MMAs, warps 5-12 generate x / split W,
// warps 13-16 drain the accumulator.  Every role is a chain of dependent instructions, so a
// warp issues one instruction every ~5 cycles: the work of a stage is spread over enough warps
// that none of them needs more than the ~800 cycles the three TF32 MMAs of a stage take.
constexpr int FWD_THREADS = 17 * 32;
constexpr int FWD_MAX_STAGES = 4;
constexpr int FWD_PREFETCH = 8;                   // L2 prefetch distance of the W stream (stages)

struct FwdArgs {
  const float* fac;
  int64_t ldf;
  const int64_t* rows;
  const float* w;          // [N][F]
  int S, Q, F;             // F = S*Q + 2
  int M, N;                // rows of the batch, output columns (<= 128)
  int splits, kb_per_split, num_kb;
  int stages;              // shared-memory / tensor-memory ring depth (<= FWD_MAX_STAGES)
  int qp, npp;             // odd pitches of the staged af (wrap-extended by 32) / sf rows
  float* partial;          // [splits][M][N]
  long long* prof;         // BSIG_CORR_PROF (profiling only): clock64 marks of CTA 0
  int dbg;                 // BSIG_CORR_DBG (profiling only): 1 no x math, 2 no W split, 4 no MMA, 8 no W copy
};

#define CORR_MARK(slot)                                                          \
  do {                                                                           \
    if (g.prof != nullptr && blockIdx.x == 0 && lane == 0) g.prof[slot] = clock64(); \
  } while (0)

__global__ void __launch_bounds__(FWD_THREADS, 1)
corr_fwd_kernel(FwdArgs g, FastDiv divQ) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // (offset arithmetic on the array keeps the shared address space: LDS / STS, not generic)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t full_bar[FWD_MAX_STAGES], conv_bar[FWD_MAX_STAGES], empty_bar[FWD_MAX_STAGES];
  __shared__ uint64_t tmem_full_bar[2], tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_slot;
  const int STAGES = g.stages;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = (g.M + BM - 1) / BM;
  const int num_items = m_tiles * g.splits;
  if (warp == 0) CORR_MARK(0);
  auto tile_b = [&](int s) { return smem + (size_t)s * 2 * TILE_BYTES; };
  auto tile_blo = [&](int s) { return tile_b(s) + TILE_BYTES; };
  float* af_s = reinterpret_cast<float*>(smem + (size_t)STAGES * 2 * TILE_BYTES);   // [128][qp]
  float* sf_s = af_s + (size_t)BM * g.qp;                                           // [128][npp]

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 128);              // one cp.async arrival per producer lane
      mbar_init(&conv_bar[s], 256);
      mbar_init(&empty_bar[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    // 2 x 128 accumulator columns + (hi 32 | lo 32) columns of generated A per stage
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(&tmem_base_slot)),
                 "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_slot;
  if (warp == 0) CORR_MARK(1);                       // set-up done

  if (warp < 4) {
    // ------------------------------------------------ producers: W tiles by 8-byte cp.async
    // W [N][F] has rows that are only 8-byte aligned (F = s*q + 2 is even, never a multiple of
    // four for the reference's windows), which TMA cannot address (16-byte box origins): each
    // producer warp copies 32 rows of the 128 x 32 tile as 8-byte pieces straight into the
    // 128B-swizzled layout the tensor core reads; completion arrives on the stage's mbarrier.
    // lane -> (row parity h, 8-byte piece pc); iteration i -> row 32*warp + 2i + h
    const int h = lane >> 4, pc = lane & 15;
    const uint32_t c0 = (uint32_t)((pc >> 1) ^ h);          // 16-byte chunk before the row swizzle
    const uint32_t dst_lane = (uint32_t)(warp * 32 + h) * 128u + (uint32_t)((pc & 1) << 3);
    const float* src_lane = g.w + (int64_t)(warp * 32 + h) * g.F + 2 * pc;
    int s = 0;
    uint32_t ph = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      const int split = item % g.splits;
      const int kb0 = split * g.kb_per_split, kb1 = min(g.num_kb, kb0 + g.kb_per_split);
      for (int kb = kb0; kb < kb1; ++kb, ph ^= (++s == STAGES), s = (s == STAGES ? 0 : s)) {
        const int k0 = kb * BK;
        // warm L2 a few stages ahead: one 128-byte line per lane (row 32*warp + lane)
        if (kb + FWD_PREFETCH < kb1 && warp * 32 + lane < g.N)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(g.w + (int64_t)(warp * 32 + lane) * g.F +
                                                       k0 + FWD_PREFETCH * BK));
        mbar_wait(&empty_bar[s], ph ^ 1);
        const uint32_t dst_base = smem_u32(tile_b(s)) + dst_lane;
        const bool k_ok = k0 + 2 * pc < g.F;
        if (!(g.dbg & 8))
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const uint32_t dst = dst_base + (uint32_t)(2 * i) * 128u + ((c0 ^ (uint32_t)((2 * i) & 7)) << 4);
          const bool valid = k_ok && (warp * 32 + 2 * i + h) < g.N;
          const float* src = valid ? src_lane + (int64_t)(2 * i) * g.F + k0 : g.w;
          asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src),
                       "r"(valid ? 8 : 0)
                       : "memory");
        }
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(
                         smem_u32(&full_bar[s]))
                     : "memory");
        if (warp == 0 && kb == kb0) CORR_MARK(2);      // first stage issued
      }
    }
    if (warp == 0) CORR_MARK(3);                     // all stages issued
  } else if (warp == 4) {
    // ------------------------------------------------ MMA issuer (A from tensor memory)
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) |
                             ((uint32_t)(BM >> 4) << 24);
      int s = 0, tcount = 0;
      uint32_t ph = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++tcount) {
        const int split = item % g.splits;
        const int kb0 = split * g.kb_per_split, kb1 = min(g.num_kb, kb0 + g.kb_per_split);
        const int buf = tcount & 1;
        mbar_wait(&tmem_empty_bar[buf], ((tcount >> 1) & 1) ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d_tmem = tmem_base + (uint32_t)(buf * BN);
        for (int kb = kb0; kb < kb1; ++kb, ph ^= (++s == STAGES), s = (s == STAGES ? 0 : s)) {
          mbar_wait(&conv_bar[s], ph);
          if (kb == kb0) CORR_MARK(4);                 // first stage converted
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          // descriptors once per stage (the issuing thread is one chain of dependent
          // instructions); a K step = 32 bytes = +2 in the address field
          const uint64_t db_hi = umma_desc(smem_u32(tile_b(s))), db_lo = umma_desc(smem_u32(tile_blo(s)));
          const uint32_t a_tm0 = tmem_base + 256u + (uint32_t)(s * 64);
          if (!(g.dbg & 4))
#pragma unroll
          for (int k = 0; k < BK / 8; ++k) {
            const uint32_t a_tm = a_tm0 + (uint32_t)(k * 8);
            const uint32_t acc = (kb > kb0 || k > 0) ? 1u : 0u;
            umma_tf32_ts(d_tmem, a_tm, db_hi + 2 * k, idesc, acc);
            umma_tf32_ts(d_tmem, a_tm, db_lo + 2 * k, idesc, 1u);
            umma_tf32_ts(d_tmem, a_tm + 32u, db_hi + 2 * k, idesc, 1u);
          }
          umma_commit(&empty_bar[s]);
        }
        umma_commit(&tmem_full_bar[buf]);
        CORR_MARK(5);                                  // last MMA issued
      }
    }
  } else if (warp < 13) {
    // ------------------------------------------------ generators: x tiles -> TMEM, W hi/lo
    // two warps per TMEM lane quadrant: each generates 16 of the 32 K columns of its rows
    const int t = threadIdx.x - 5 * 32;         // 0..255
    const int half = (warp - 5) >> 2;           // K columns [16*half, 16*half + 16) of a block
    const int m = (warp & 3) * 32 + lane;       // row of the tile = TMEM lane of this thread
    const uint32_t SQ = (uint32_t)(g.F - 2);
    int s = 0;
    uint32_t ph = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      const int mt = item / g.splits, split = item - mt * g.splits;
      const int kb0 = split * g.kb_per_split, kb1 = min(g.num_kb, kb0 + g.kb_per_split);
      const int m0 = mt * BM;
      const uint32_t p_base = fdiv((uint32_t)(kb0 * BK), divQ);
      const uint32_t p_last = min((uint32_t)g.S - 1, fdiv((uint32_t)(min(kb1 * BK, g.F) - 1), divQ));
      const int np = p_base < (uint32_t)g.S ? (int)(p_last - p_base) + 1 : 0;
      // stage the factors of this item's rows: af [128][Q + 32] (the first 32 entries repeated
      // behind the row, so that a block of 32 consecutive q never has to wrap), the sf columns
      // of the K slice.  Thread pair (t, t+128) shares row t & 127; 4-byte cp.async each.
      asm volatile("bar.sync 1, 256;" ::: "memory");     // previous item fully generated
      {
        // generator warp gw copies rows gw, gw+8, ... (16 rows): lanes run along the feature
        // index, so every 4-byte cp.async instruction reads one 128-byte line of a factor row;
        // the 16 source-row indices are fetched by 16 lanes at once (one load latency)
        const int gw = warp - 5;
        const int qx = g.Q + 32;
        int64_t my_src = -1;
        if (lane < 16 && m0 + gw + 8 * lane < g.M)
          my_src = g.rows ? __ldg(g.rows + m0 + gw + 8 * lane) : (int64_t)(m0 + gw + 8 * lane);
        // (pointer-increment loops: this warp is a single chain of dependent instructions, every
        // instruction of address arithmetic costs ~5 cycles)
        const int n_full = g.Q >> 5;                       // whole 32-float groups of af
        const int q_tail = (n_full << 5) + lane;
        const int ext_src = lane < g.Q ? lane : 0;         // (Q < 32: extension unused)
#pragma unroll 2
        for (int u = 0; u < 16; ++u) {
          const int64_t src = __shfl_sync(0xffffffffu, my_src, u);
          const int r = gw + 8 * u;
          const int sz = src >= 0 ? 4 : 0;               // rows past the batch: zero fill
          const float* frow = g.fac + (src >= 0 ? src : 0) * g.ldf;
          const float* ap = frow + g.S + lane;
          uint32_t ad = smem_u32(af_s + r * g.qp) + 4u * lane;
          for (int i = 0; i < n_full; ++i, ap += 32, ad += 128u)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(ad), "l"(ap), "r"(sz)
                         : "memory");
          if (q_tail < g.Q)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(ad), "l"(ap), "r"(sz)
                         : "memory");
          // [Q, Q+32): the first entries again; then mean, std
          const uint32_t xd = smem_u32(af_s + r * g.qp + g.Q) + 4u * lane;
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(xd),
                       "l"(frow + g.S + ext_src), "r"(sz)
                       : "memory");
          if (lane < 2)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(xd + 128u),
                         "l"(frow + g.S + g.Q + lane), "r"(sz)
                         : "memory");
          const float* sp = frow + p_base + lane;
          uint32_t sd_ = smem_u32(sf_s + r * g.npp) + 4u * lane;
          for (int c = lane; c < np; c += 32, sp += 32, sd_ += 128u)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(sd_), "l"(sp), "r"(sz)
                         : "memory");
        }
      }
      asm volatile("cp.async.wait_all;" ::: "memory");
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (warp == 5) CORR_MARK(6);                   // factors staged
      const float* afr = af_s + m * g.qp;
      const float mu = afr[g.Q + 32], sd = afr[g.Q + 33];
      const float* sfr = sf_s + m * g.npp;
      for (int kb = kb0; kb < kb1; ++kb, ph ^= (++s == STAGES), s = (s == STAGES ? 0 : s)) {
        // the x tile does not depend on the W data, but its TMEM slot is only free once the
        // MMAs of the previous round have retired -- which is what let the producers refill
        // the stage, i.e. what full_bar certifies
        mbar_wait(&full_bar[s], ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t k0 = (uint32_t)kb * BK + 16u * (uint32_t)half;
        const uint32_t pcur = fdiv(k0, divQ);
        const uint32_t q0 = k0 - pcur * (uint32_t)g.Q;
        uint32_t hi[16], lo[16];
        if (g.dbg & 1) {
#pragma unroll
          for (int i = 0; i < 16; ++i) hi[i] = lo[i] = 0u;
        } else if (g.Q >= 32 && k0 + 16 <= SQ) {
          // common case: 16 products from at most two state features, no wrap in af
          const uint32_t n1 = (uint32_t)g.Q - q0;          // elements left in feature row pcur
          const float sv0 = sfr[pcur - p_base];
          const float sv1 = n1 < 16 ? sfr[pcur + 1 - p_base] : 0.f;
          const float* ap = afr + q0;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float xv = ((uint32_t)i < n1 ? sv0 : sv1) * ap[i];
            const uint32_t hb = __float_as_uint(xv) & 0xffffe000u;
            hi[i] = hb;
            lo[i] = __float_as_uint(xv - __uint_as_float(hb));
          }
        } else {
          // narrow action windows (several wraps per block) and the tail block with the two
          // statistics / the zero padding behind them
          uint32_t pp = pcur, q = q0;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float xv;
            if (pp < (uint32_t)g.S) {
              xv = sfr[pp - p_base] * afr[q];
            } else {
              const uint32_t tail = k0 + i - SQ;
              xv = tail == 0 ? mu : (tail == 1 ? sd : 0.f);
            }
            const uint32_t hb = __float_as_uint(xv) & 0xffffe000u;
            hi[i] = hb;
            lo[i] = __float_as_uint(xv - __uint_as_float(hb));
            if (++q == (uint32_t)g.Q) {
              q = 0;
              ++pp;
            }
          }
        }
        const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + 256u +
                               (uint32_t)(s * 64 + 16 * half);
        BSIG_TMEM_ST16(taddr, hi);
        BSIG_TMEM_ST16(taddr + 32u, lo);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        // W tile: hi in place, lo beside it
        float4* b = reinterpret_cast<float4*>(tile_b(s));
        float4* blo = reinterpret_cast<float4*>(tile_blo(s));
        if (!(g.dbg & 2))
#pragma unroll
        for (int u = 0; u < TILE_BYTES / 16 / 256; ++u) {
          const int idx = t + 256 * u;
          const float4 v = b[idx];
          float4 hh, l;
          hh.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u);
          hh.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
          hh.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u);
          hh.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
          l.x = v.x - hh.x; l.y = v.y - hh.y; l.z = v.z - hh.z; l.w = v.w - hh.w;
          b[idx] = hh;
          blo[idx] = l;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        mbar_arrive(&conv_bar[s]);
      }
    }
  } else {
    // ------------------------------------------------ epilogue: raw partial accumulators
    const int qd = warp & 3;
    int tcount = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++tcount) {
      const int mt = item / g.splits, split = item - mt * g.splits;
      const int buf = tcount & 1;
      mbar_wait(&tmem_full_bar[buf], (tcount >> 1) & 1);
      if (warp == 13) CORR_MARK(7);                  // accumulator complete
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int row = mt * BM + qd * 32 + lane;
      const bool row_ok = row < g.M;
      // partial sums are stored TRANSPOSED, [split][j][m]: the lanes of a warp hold 32
      // consecutive rows m, so every store instruction writes one 128-byte line
      float* pcol = g.partial + (int64_t)split * g.N * g.M + (row_ok ? row : 0);
#pragma unroll 1
      for (int cb = 0; cb < BN / 32; ++cb) {
        if (cb * 32 >= g.N) break;
        uint32_t r[32];
        const uint32_t taddr =
            tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(buf * BN + cb * 32);
        BSIG_TMEM_LD32(r, taddr);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (!row_ok) continue;
        float* pp = pcol + (int64_t)(cb * 32) * g.M;
        if (cb * 32 + 32 <= g.N) {                       // full block: no per-element predicate
#pragma unroll
          for (int c = 0; c < 32; ++c, pp += g.M) *pp = __uint_as_float(r[c]);
        } else {
#pragma unroll
          for (int c = 0; c < 32; ++c, pp += g.M)
            if (cb * 32 + c < g.N) *pp = __uint_as_float(r[c]);
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(&tmem_empty_bar[buf]);
      if (warp == 13) CORR_MARK(8);                  // partials stored
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) CORR_MARK(9);
  if (warp == 4) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512)
                 : "memory");
  }
}

// y[m][j] = act( sum_s partial[s][j][m] + b[j] ): fixed-order, four threads per element each
// summing a quarter of the splits (eight loads in flight), combined through shared memory in
// quarter order -- deterministic, and 4x the memory-level parallelism of one thread per element.
// act: 0 none, 1 tanh, 2 random Fourier features: y[m][j] = scale cos(sum), y[m][N + j] =
// scale sin(sum) (models/rff.py:128-132; full-precision sincosf, as the SIMT engine).
__global__ void __launch_bounds__(256)
corr_reduce_kernel(const float* __restrict__ partial, const float* __restrict__ bias,
                   float* __restrict__ y, int M, int N, int splits, int act, float scale) {
  __shared__ float part[4][64];
  const int el = threadIdx.x & 63, grp = threadIdx.x >> 6;
  const int64_t total = (int64_t)M * N;
  const int64_t e = (int64_t)blockIdx.x * 64 + el;            // e = j * M + m
  const int per = (splits + 3) / 4;
  const int s0 = grp * per, s1 = min(splits, s0 + per);
  float acc = 0.f;
  if (e < total) {
    int sidx = s0;
    for (; sidx + 8 <= s1; sidx += 8) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = __ldcg(partial + (int64_t)(sidx + u) * total + e);
#pragma unroll
      for (int u = 0; u < 8; ++u) acc += v[u];
    }
    for (; sidx < s1; ++sidx) acc += __ldcg(partial + (int64_t)sidx * total + e);
  }
  part[grp][el] = acc;
  __syncthreads();
  if (grp == 0 && e < total) {
    const float sum = ((part[0][el] + part[1][el]) + part[2][el]) + part[3][el];
    const int j = (int)(e / M), m = (int)(e - (int64_t)j * M);
    float v = sum + (bias != nullptr ? __ldg(bias + j) : 0.f);
    if (act == 2) {
      float sn, co;
      sincosf(v, &sn, &co);
      y[(int64_t)m * 2 * N + j] = scale * co;
      y[(int64_t)m * 2 * N + N + j] = scale * sn;
    } else {
      if (act == 1) v = tanhf(v);
      y[(int64_t)m * N + j] = v;
    }
  }
}

// ------------------------------------------------------------------ weight gradient (+ Adam)
constexpr int WG_TN = 64;                         // k columns per output tile
constexpr int WG_EPI_WARPS = 16;
// warp 0 MMA, warps 1-4 generate, warp 5 loads state-feature columns, warps 6-21 epilogue
constexpr int WG_THREADS = (6 + WG_EPI_WARPS) * 32;
constexpr int WG_STAGE_BYTES = 2 * 4 * WG_TN * 128;   // hi + lo, four [64 x 32] sub-tiles = 64 KB
constexpr int WG_EP = 65;                         // pitch of the epilogue staging tile

struct WgradArgs {
  const float* dy;         // [M][N]  (N = n_out, contiguous)
  const float* fac;
  int64_t ldf;
  const int64_t* rows;
  int S, Q, F;
  int M, N;                // batch rows (<= 128), n_out (<= 128)
  int KP;                  // M rounded up to 8
  int NP;                  // pitch of the transposed factor rows in shared memory
  int npt;                 // sf columns one tile of WG_TN summary columns can touch
  int num_tiles;
  float* dw;               // nullable: store the gradient [N][F]
  float* w;                // Adam in the epilogue when exp_avg != nullptr
  float* exp_avg;
  float* exp_avg_sq;
  float one_minus_b1, b2, one_minus_b2, step_size, inv_bc2_sqrt, eps, gscale;
};

__global__ void __launch_bounds__(WG_THREADS, 1) corr_wgrad_kernel(WgradArgs g, FastDiv divQ) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t b_full[2], b_empty[2], acc_full[2], acc_empty[2], sf_full[2];
  __shared__ uint32_t tmem_base_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* stage_b = smem;                                               // 2 x 64 KB
  float* ep_s = reinterpret_cast<float*>(smem + 2 * WG_STAGE_BYTES);      // [128][65]
  float* afT = ep_s + 128 * WG_EP;                                        // [Q][NP]
  float* sfc = afT + (size_t)g.Q * g.NP;                                  // [2][npt][NP]
  float* mu_s = sfc + (size_t)2 * g.npt * g.NP;                           // [128]
  float* sd_s = mu_s + 128;                                               // [128]
  int64_t* src_s = reinterpret_cast<int64_t*>(sd_s + 128);                // [128] factor row of n

  // tiles are dealt round-robin: at any moment the CTAs work on NEIGHBOURING 256-byte segments
  // of every weight / moment row (148 x 256 B = 37 KB contiguous per row), which is what keeps
  // the DRAM pages open; a contiguous range per CTA scatters 57 k concurrent segments over as
  // many pages (measured: 3.7 TB/s)
  const uint32_t SQ = (uint32_t)(g.F - 2);

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&b_full[i], 128);
      mbar_init(&b_empty[i], 1);
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], WG_EPI_WARPS * 32);
      mbar_init(&sf_full[i], 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    // columns [0,256): dy^T as tf32 hi (0..) and lo (128..); [256, 384): two accumulators
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(&tmem_base_slot)),
                 "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_slot;

  if (warp >= 6 && warp < 10) {
    // ---- one-off (epilogue warps 6-9, one per TMEM lane quadrant): dy^T -> tensor memory
    //      (lane = output j, column = batch row n)
    const int j = (warp & 3) * 32 + lane;        // TMEM lane of this thread
    const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    for (int n0 = 0; n0 < g.KP; n0 += 32) {
      float v[32];
#pragma unroll
      for (int u = 0; u < 32; ++u) {             // 32 independent loads in flight
        const int n = n0 + u;
        v[u] = (n < g.M && j < g.N) ? __ldg(g.dy + (int64_t)n * g.N + j) : 0.f;
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (n0 + 8 * c < g.KP) {                 // warp-uniform
          uint32_t hi[8], lo[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const uint32_t h = __float_as_uint(v[8 * c + u]) & 0xffffe000u;
            hi[u] = h;
            lo[u] = __float_as_uint(v[8 * c + u] - __uint_as_float(h));
          }
          BSIG_TMEM_ST8(lane_addr + (uint32_t)(n0 + 8 * c), hi);
          BSIG_TMEM_ST8(lane_addr + 128u + (uint32_t)(n0 + 8 * c), lo);
        }
      }
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  } else if (warp >= 10 && warp < 14) {
    // ---- one-off (epilogue warps 10-13): warp gw stages batch rows gw, gw+4, ... transposed
    //      ([feature][row]); lanes run along the feature index so that every 4-byte cp.async
    //      instruction reads one line
    const int t = threadIdx.x - 10 * 32;         // 0..127
    {
      const int gw = warp - 10;
      const uint32_t pitch = 4u * (uint32_t)g.NP;
      // lane u holds the source row of batch row gw + 4u (KP <= 128: 32 rows per warp)
      int64_t my_src = -1;
      if (gw + 4 * lane < g.M) my_src = g.rows ? __ldg(g.rows + gw + 4 * lane) : (int64_t)(gw + 4 * lane);
      const int n_rows = (g.KP - gw + 3) >> 2;
      for (int u = 0; u < n_rows; ++u) {
        const int64_t src = __shfl_sync(0xffffffffu, my_src, u);
        const int n = gw + 4 * u;
        const int sz = src >= 0 ? 4 : 0;
        const float* frow = g.fac + (src >= 0 ? src : 0) * g.ldf;
        const float* ap = frow + g.S + lane;
        uint32_t ad = smem_u32(afT + n) + pitch * lane;
        for (int q = lane; q < g.Q; q += 32, ap += 32, ad += 32u * pitch)
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(ad), "l"(ap), "r"(sz)
                       : "memory");
        if (lane == 0) src_s[n] = src;             // (-1: row of the zero padding)
        if (lane < 2)
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(
                           smem_u32((lane == 0 ? mu_s : sd_s) + n)),
                       "l"(frow + g.S + g.Q + lane), "r"(sz)
                       : "memory");
      }
      if (t >= g.KP) {
        mu_s[t] = 0.f;
        sd_s[t] = 0.f;
        src_s[t] = -1;
      }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
  }
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  const int n_chunks = g.KP / 4;                 // 16-byte chunks (4 batch rows) per k row
  if (warp == 0) {
    // ------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(WG_TN >> 3) << 17) |
                             ((uint32_t)(128 >> 4) << 24);
      int it = 0;
      for (int tile = blockIdx.x; tile < g.num_tiles; tile += gridDim.x, ++it) {
        const int s = it & 1;
        const uint32_t par = (it >> 1) & 1;
        mbar_wait(&acc_empty[s], par ^ 1);
        mbar_wait(&b_full[s], par);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d_tmem = tmem_base + 256u + (uint32_t)(s * WG_TN);
        const uint32_t b_hi = smem_u32(stage_b + (size_t)s * WG_STAGE_BYTES);
        const uint32_t b_lo = b_hi + WG_STAGE_BYTES / 2;
        const uint64_t db_hi = umma_desc(b_hi), db_lo = umma_desc(b_lo);
        for (int ks = 0; ks < g.KP / 8; ++ks) {
          // sub-tile (32 batch rows) = 8 KB = +512 descriptor units, K step inside it = +2
          const uint64_t off = (uint64_t)((ks >> 2) * (WG_TN * 128 / 16) + (ks & 3) * 2);
          const uint32_t a_hi = tmem_base + (uint32_t)(ks * 8), a_lo = a_hi + 128u;
          umma_tf32_ts(d_tmem, a_hi, db_hi + off, idesc, ks > 0 ? 1u : 0u);
          umma_tf32_ts(d_tmem, a_hi, db_lo + off, idesc, 1u);
          umma_tf32_ts(d_tmem, a_lo, db_hi + off, idesc, 1u);
        }
        umma_commit(&b_empty[s]);
        umma_commit(&acc_full[s]);
      }
    }
  } else if (warp <= 4) {
    // ------------------------------------------------ generators: x^T tiles, K-major, SW128
    const int t = threadIdx.x - 32;              // 0..127
    const int r = t & (WG_TN - 1);               // k row of the tile owned by this thread
    int it = 0;
    for (int tile = blockIdx.x; tile < g.num_tiles; tile += gridDim.x, ++it) {
      const int s = it & 1;
      mbar_wait(&b_empty[s], ((it >> 1) & 1) ^ 1);
      mbar_wait(&sf_full[s], (it >> 1) & 1);
      const uint32_t p_base = min(fdiv(min((uint32_t)tile * WG_TN, SQ - 1), divQ), (uint32_t)g.S - 1);
      const float* sfT = sfc + (size_t)s * g.npt * g.NP;
      uint8_t* hi_base = stage_b + (size_t)s * WG_STAGE_BYTES;
      uint8_t* lo_base = hi_base + WG_STAGE_BYTES / 2;
      const uint32_t k = (uint32_t)tile * WG_TN + (uint32_t)r;
      const uint32_t pk = fdiv(min(k, SQ), divQ);
      const uint32_t qk = k - pk * (uint32_t)g.Q;
      const bool prod = k < SQ;
      const float* sfp = sfT + (size_t)(prod ? pk - p_base : 0) * g.NP;
      const float* afp = afT + (size_t)(prod ? qk : 0) * g.NP;
      const float* tailp = k == SQ ? mu_s : sd_s;
      const bool tail = (k == SQ) || (k == SQ + 1);
      for (int c = t >> 6; c < n_chunks; c += 2) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (prod) {
          const float4 a = *reinterpret_cast<const float4*>(afp + 4 * c);
          const float4 sv = *reinterpret_cast<const float4*>(sfp + 4 * c);
          v.x = sv.x * a.x; v.y = sv.y * a.y; v.z = sv.z * a.z; v.w = sv.w * a.w;
        } else if (tail) {
          v = *reinterpret_cast<const float4*>(tailp + 4 * c);
        }
        float4 h, l;
        h.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u);
        h.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
        h.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u);
        h.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
        l.x = v.x - h.x; l.y = v.y - h.y; l.z = v.z - h.z; l.w = v.w - h.w;
        const uint32_t off = (uint32_t)(c >> 3) * (WG_TN * 128) + (uint32_t)r * 128 +
                             (uint32_t)(((c & 7) ^ (r & 7)) << 4);
        *reinterpret_cast<float4*>(hi_base + off) = h;
        *reinterpret_cast<float4*>(lo_base + off) = l;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_arrive(&b_full[s]);
    }
  } else if (warp == 5) {
    // ------------------------------------------------ loader: the (at most npt) state-feature
    // columns a tile touches, [column][batch row], straight from the L2-resident factor rows;
    // runs up to two tiles ahead of the generators (its buffer is free as soon as they have
    // finished the tile that used it: their own arrival on b_full)
    int it = 0;
    for (int tile = blockIdx.x; tile < g.num_tiles; tile += gridDim.x, ++it) {
      const int s = it & 1;
      if (it >= 2) mbar_wait(&b_full[s], ((it >> 1) & 1) ^ 1);
      const uint32_t p_base = min(fdiv(min((uint32_t)tile * WG_TN, SQ - 1), divQ), (uint32_t)g.S - 1);
      float* dst = sfc + (size_t)s * g.npt * g.NP;
      // eight loads in flight per lane (the shared-memory stores would otherwise serialise
      // them: the compiler must assume dst aliases src_s)
      const int per_col = (g.KP + 31) >> 5, total = g.npt * per_col;
      for (int e0 = 0; e0 < total; e0 += 8) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int e = e0 + u;
          const int c = e / per_col, n = (e - c * per_col) * 32 + lane;
          v[u] = 0.f;
          if (e < total && n < g.KP && p_base + c < (uint32_t)g.S) {
            const int64_t src = src_s[n];
            if (src >= 0) v[u] = __ldg(g.fac + src * g.ldf + p_base + c);
          }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int e = e0 + u;
          const int c = e / per_col, n = (e - c * per_col) * 32 + lane;
          if (e < total && n < g.KP) dst[c * g.NP + n] = v[u];
        }
      }
      mbar_arrive(&sf_full[s]);
    }
  } else if (warp >= 6) {
    // ------------------------------------------------ epilogue: Adam / gradient store
    // 16 warps: four per TMEM lane quadrant, each drains 16 of the 64 accumulator columns into
    // the staging tile; then warp ew owns rows ew, ew+16, ... (8 rows) and lane l the column
    // pair 2l: every row is one coalesced 256-byte access per array.  The HBM stream (read and
    // write W, exp_avg, exp_avg_sq: 24 B per parameter) is the whole cost of the kernel.
    const int ew = warp - 6;                      // 0..15
    const int qd = warp & 3;                      // TMEM lane quadrant
    const int part = ew >> 2;                     // which 16 of the 64 accumulator columns
    const bool adam = g.exp_avg != nullptr;
    // Software pipeline over (tile, sub-batch of 4 rows): the loads of the NEXT sub-batch are in
    // flight while Adam runs on the current one (two register sets of 4 rows x 3 arrays), across
    // tile boundaries as well -- without it the kernel alternated between a phase that saturates
    // HBM (loads of tile t + stores of tile t-1) and a phase that leaves it idle (the
    // arithmetic), 4.5 + 2.8 us per tile.
    struct RowSet {
      float2 w[4], m[4], v[4];
    };
    const int64_t step8 = 16 * (int64_t)g.F;      // row stride of a warp: 16 rows
    auto issue = [&](RowSet& rs, int tile, int q) {
      const int64_t kcol = (int64_t)tile * WG_TN + 2 * lane;
      if (!adam || kcol >= g.F) return;
      const int64_t e0 = (int64_t)(ew + 64 * q) * g.F + kcol;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (ew + 16 * (4 * q + u) < g.N) {
          rs.w[u] = *reinterpret_cast<const float2*>(g.w + e0 + u * step8);
          rs.m[u] = *reinterpret_cast<const float2*>(g.exp_avg + e0 + u * step8);
          rs.v[u] = *reinterpret_cast<const float2*>(g.exp_avg_sq + e0 + u * step8);
        }
      }
    };
    // (SFU square root and reciprocal, <= 2 ulp each: the IEEE sequences of sqrtf and the
    // division were half of the epilogue's instructions; the difference to bsig_adam_step's
    // arithmetic is ~2e-7 of an lr-sized step)
    auto upd = [&](float& pp, float gg, float& mm, float& vv) {
      gg *= g.gscale;
      mm = mm + (gg - mm) * g.one_minus_b1;
      vv = vv * g.b2 + g.one_minus_b2 * gg * gg;
      float sq;
      asm("sqrt.approx.f32 %0, %1;" : "=f"(sq) : "f"(vv));
      const float denom = sq * g.inv_bc2_sqrt + g.eps;
      pp = pp - g.step_size * __fdividef(mm, denom);
    };
    auto apply = [&](RowSet& rs, int tile, int q) {
      const int64_t kcol = (int64_t)tile * WG_TN + 2 * lane;
      if (kcol >= g.F) return;
      const int64_t e0 = (int64_t)(ew + 64 * q) * g.F + kcol;
      const float* gs = ep_s + (ew + 64 * q) * WG_EP + 2 * lane;
#pragma unroll
      for (int u = 0; u < 4; ++u, gs += 16 * WG_EP) {
        if (ew + 16 * (4 * q + u) < g.N) {
          const float g0 = gs[0], g1 = gs[1];
          const int64_t e = e0 + u * step8;
          if (adam) {
            upd(rs.w[u].x, g0, rs.m[u].x, rs.v[u].x);
            upd(rs.w[u].y, g1, rs.m[u].y, rs.v[u].y);
            *reinterpret_cast<float2*>(g.w + e) = rs.w[u];
            *reinterpret_cast<float2*>(g.exp_avg + e) = rs.m[u];
            *reinterpret_cast<float2*>(g.exp_avg_sq + e) = rs.v[u];
          }
          if (g.dw != nullptr) *reinterpret_cast<float2*>(g.dw + e) = make_float2(g0, g1);
        }
      }
    };
    RowSet ra, rb;
    if ((int)blockIdx.x < g.num_tiles) issue(ra, blockIdx.x, 0);
    int it = 0;
    for (int tile = blockIdx.x; tile < g.num_tiles; tile += gridDim.x, ++it) {
      const int s = it & 1;
      mbar_wait(&acc_full[s], (it >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      {
        uint32_t r[16];
        const uint32_t taddr = tmem_base + ((uint32_t)(qd * 32) << 16) + 256u +
                               (uint32_t)(s * WG_TN + part * 16);
        BSIG_TMEM_LD16(r, taddr);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        mbar_arrive(&acc_empty[s]);
        asm volatile("bar.sync 1, 512;" ::: "memory");    // previous tile's staging fully consumed
        float* erow = ep_s + (qd * 32 + lane) * WG_EP + part * 16;
#pragma unroll
        for (int c = 0; c < 16; ++c) erow[c] = __uint_as_float(r[c]);
      }
      asm volatile("bar.sync 1, 512;" ::: "memory");
      issue(rb, tile, 1);
      apply(ra, tile, 0);
      if (tile + (int)gridDim.x < g.num_tiles) issue(ra, tile + gridDim.x, 0);
      apply(rb, tile, 1);
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512)
                 : "memory");
  }
}

// ------------------------------------------------------------------ host side
struct FwdPlan {
  bool ok;
  int splits, kb_per_split, num_kb, stages, qp, npp;
  size_t smem;
};

static FwdPlan plan_fwd(int64_t m, int64_t n_out, int64_t s, int64_t q) {
  FwdPlan p = {};
  const int64_t F = s * q + 2;
  // (F odd -- Humanoid: 5*107 x 5*21 + 2 -- leaves the rows of W only 4-byte aligned: the 8-byte
  // copies / float2 accesses of these kernels need an even F)
  if (m < 1 || n_out < 1 || n_out > 128 || s < 1 || q < 1 || F >= (1 << 20) || (F & 1)) return p;
  const int64_t m_tiles = ceil_div(m, BM);
  p.num_kb = (int)ceil_div(F, BK);
  int64_t splits = std::max<int64_t>(1, std::min<int64_t>(p.num_kb, sm_count() / m_tiles));
  p.kb_per_split = (int)ceil_div(p.num_kb, splits);
  p.splits = (int)ceil_div(p.num_kb, p.kb_per_split);
  p.qp = (int)((q + 34) | 1);          // af + 32 wrap entries + mean, std
  const int64_t np = ((int64_t)p.kb_per_split * BK) / q + 2;
  p.npp = (int)(np | 1);
  const size_t factors = (size_t)BM * (p.qp + p.npp) * 4;
  for (int st = FWD_MAX_STAGES; st >= 2; --st) {
    const size_t need = (size_t)st * 2 * TILE_BYTES + factors + 1024;
    if (need <= 220 * 1024) {
      p.stages = st;
      p.smem = need;
      p.ok = true;
      break;
    }
  }
  return p;
}

struct WgPlan {
  bool ok;
  int KP, NP, npt, num_tiles, grid;
  size_t smem;
};

static WgPlan plan_wgrad(int64_t m, int64_t n_out, int64_t s, int64_t q) {
  WgPlan p = {};
  const int64_t F = s * q + 2;
  if (m < 1 || m > 128 || n_out < 1 || n_out > 128 || s < 1 || q < 1 || F >= (1 << 20) || (F & 1)) return p;
  p.KP = (int)(ceil_div(m, 8) * 8);
  p.NP = p.KP + 4;
  p.num_tiles = (int)ceil_div(F, WG_TN);
  p.grid = (int)std::min<int64_t>(p.num_tiles, sm_count());
  p.npt = (int)((WG_TN - 1) / q + 2);
  p.smem = (size_t)2 * WG_STAGE_BYTES + (size_t)128 * WG_EP * 4 +
           (size_t)(q + 2 * p.npt) * p.NP * 4 + 2 * 128 * 4 + 128 * 8 + 1024;
  p.ok = p.smem <= 225 * 1024;
  return p;
}

}  // namespace corr
}  // namespace bsig

using namespace bsig;

extern "C" int bsig_corr_factors(const float* states, const float* actions, float* fac, int64_t ldf,
                                 int64_t n, int64_t t_states, int64_t t_actions, int64_t d,
                                 int64_t a, int64_t w, int use_state_diff, int time_major,
                                 int* nonfinite_flag, void* stream) {
  BSIG_REQUIRE(n >= 0 && d >= 2 && a >= 1 && w >= 1, "corr_factors: need d>=2, a>=1, w>=1");
  BSIG_REQUIRE(t_states >= w && t_actions >= w, "corr_factors: trajectories shorter than w");
  BSIG_REQUIRE(nonfinite_flag != nullptr, "corr_factors: flag pointer required");
  const int64_t pn = w * (d - 1), qn = w * a;
  BSIG_REQUIRE(ldf >= pn + qn + 2, "corr_factors: row pitch too small");
  BSIG_REQUIRE(pn < (1 << 20) && qn < (1 << 20), "corr_factors: window too wide");
  if (n == 0) return 0;
  corr::FactorArgs p;
  p.states = states; p.actions = actions; p.fac = fac; p.ldf = ldf; p.flag = nonfinite_flag;
  p.n = n;
  if (time_major) {
    p.s_stride = d; p.a_stride = a; p.s_tstride = n * d; p.a_tstride = n * a;
  } else {
    p.s_stride = t_states * d; p.a_stride = t_actions * a; p.s_tstride = d; p.a_tstride = a;
  }
  p.D = (int)d; p.A = (int)a; p.Pn = (int)pn; p.Qn = (int)qn; p.use_diff = use_state_diff;
  const size_t smem = (size_t)8 * pn * 4;
  BSIG_REQUIRE(smem <= 200 * 1024, "corr_factors: window too large for shared memory");
  if (smem > 48 * 1024)
    BSIG_CUDA(cudaFuncSetAttribute(corr::corr_factors_kernel,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  corr::corr_factors_kernel<<<(unsigned)ceil_div(n, 8), 256, smem, (cudaStream_t)stream>>>(
      p, corr::mk_div((uint64_t)(d - 1)), corr::mk_div((uint64_t)a));
  BSIG_LAUNCH_CHECK();
  return 0;
}

extern "C" int bsig_corr_linear_applicable(int64_t batch, int64_t rows_max, int64_t n_out,
                                           int64_t s, int64_t q) {
  return (corr::plan_fwd(rows_max, n_out, s, q).ok && corr::plan_fwd(batch, n_out, s, q).ok &&
          corr::plan_wgrad(batch, n_out, s, q).ok) ? 1 : 0;
}

extern "C" int64_t bsig_corr_linear_ws_bytes(int64_t m, int64_t n_out, int64_t s, int64_t q) {
  const corr::FwdPlan p = corr::plan_fwd(m, n_out, s, q);
  if (!p.ok) return 0;
  return (int64_t)p.splits * m * n_out * 4 + 1024;
}

static int corr_fwd_impl(const float* fac, int64_t ldf, const int64_t* rows, int64_t s, int64_t q,
                         const float* w, const float* b, float* y, int64_t m, int64_t n_out,
                         int act, float scale, void* ws, int64_t ws_bytes, void* stream) {
  using namespace corr;
  const FwdPlan p = plan_fwd(m, n_out, s, q);
  BSIG_REQUIRE(p.ok, "corr_linear_fwd: shape outside the fused kernel's envelope "
               "(n_out <= 128, s*q+2 even and < 2^20, factors must fit shared memory)");
  BSIG_REQUIRE(act == 0 || act == 1 || act == 2, "corr_linear_fwd: unknown activation");
  BSIG_REQUIRE(!(b == nullptr && act == 1), "corr_linear_fwd: tanh needs a bias");
  BSIG_REQUIRE((reinterpret_cast<uintptr_t>(w) & 7) == 0, "corr_linear_fwd: weight must be 8-byte aligned");
  BSIG_REQUIRE(ws != nullptr && ws_bytes >= bsig_corr_linear_ws_bytes(m, n_out, s, q),
               "corr_linear_fwd: workspace too small");
  const int64_t F = s * q + 2;
  FwdArgs g;
  g.fac = fac; g.ldf = ldf; g.rows = rows; g.w = w; g.S = (int)s; g.Q = (int)q; g.F = (int)F;
  g.M = (int)m; g.N = (int)n_out;
  g.splits = p.splits; g.kb_per_split = p.kb_per_split; g.num_kb = p.num_kb; g.stages = p.stages;
  g.qp = p.qp; g.npp = p.npp;
  { const char* e = getenv("BSIG_CORR_DBG"); g.dbg = e ? atoi(e) : 0; }
  static long long* prof_dev = nullptr;
  g.prof = nullptr;
  if (getenv("BSIG_CORR_PROF") != nullptr) {
    if (prof_dev == nullptr) BSIG_CUDA(cudaMalloc(&prof_dev, 16 * sizeof(long long)));
    BSIG_CUDA(cudaMemset(prof_dev, 0, 16 * sizeof(long long)));
    g.prof = prof_dev;
  }
  g.partial = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255);
  const int64_t items = ceil_div(m, BM) * p.splits;
  BSIG_CUDA(cudaFuncSetAttribute(corr_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)p.smem));
  corr_fwd_kernel<<<(unsigned)std::min<int64_t>(items, sm_count()), FWD_THREADS, p.smem,
                    (cudaStream_t)stream>>>(g, mk_div((uint64_t)q));
  BSIG_LAUNCH_CHECK();
  if (g.prof != nullptr) {
    long long h[16];
    BSIG_CUDA(cudaDeviceSynchronize());
    BSIG_CUDA(cudaMemcpy(h, prof_dev, sizeof(h), cudaMemcpyDeviceToHost));
    fprintf(stderr, "corr_fwd CTA0 cycles since start: setup %lld  first-issue %lld  all-issued %lld  "
            "staged %lld  first-converted %lld  last-mma %lld  acc-complete %lld  stored %lld  end %lld\n",
            h[1] - h[0], h[2] - h[0], h[3] - h[0], h[6] - h[0], h[4] - h[0], h[5] - h[0], h[7] - h[0],
            h[8] - h[0], h[9] - h[0]);
  }
  const int64_t total = m * n_out;
  corr_reduce_kernel<<<(unsigned)ceil_div(total, 64), 256, 0, (cudaStream_t)stream>>>(
      g.partial, b, y, (int)m, (int)n_out, p.splits, act, scale);
  BSIG_LAUNCH_CHECK();
  return 0;
}

extern "C" int bsig_corr_linear_fwd(const float* fac, int64_t ldf, const int64_t* rows, int64_t s,
                                    int64_t q, const float* w, const float* b, float* y, int64_t m,
                                    int64_t n_out, int act, void* ws, int64_t ws_bytes,
                                    void* stream) {
  BSIG_REQUIRE(act == BSIG_ACT_NONE || act == BSIG_ACT_TANH, "corr_linear_fwd: unknown activation");
  return corr_fwd_impl(fac, ldf, rows, s, q, w, b, y, m, n_out, act == BSIG_ACT_TANH ? 1 : 0, 1.f, ws,
                       ws_bytes, stream);
}

extern "C" int bsig_corr_rff_features(const float* fac, int64_t ldf, const int64_t* rows, int64_t s,
                                      int64_t q, const float* coeff, float* out, int64_t m,
                                      int64_t nf_half, float scale, void* ws, int64_t ws_bytes,
                                      void* stream) {
  return corr_fwd_impl(fac, ldf, rows, s, q, coeff, nullptr, out, m, nf_half, 2, scale, ws, ws_bytes,
                       stream);
}

extern "C" int bsig_corr_linear_wgrad(const float* dy, const float* fac, int64_t ldf,
                                      const int64_t* rows, int64_t s, int64_t q, int64_t m,
                                      int64_t n_out, float* dw, float* w, float* exp_avg,
                                      float* exp_avg_sq, int64_t step, float lr, float beta1,
                                      float beta2, float eps, float grad_scale, void* stream) {
  using namespace corr;
  const WgPlan p = plan_wgrad(m, n_out, s, q);
  BSIG_REQUIRE(p.ok, "corr_linear_wgrad: shape outside the fused kernel's envelope "
               "(batch <= 128, n_out <= 128, s*q+2 even and < 2^20, factors must fit shared memory)");
  const bool adam = exp_avg != nullptr;
  BSIG_REQUIRE(adam || dw != nullptr, "corr_linear_wgrad: nothing to do (no dw, no Adam state)");
  BSIG_REQUIRE(!adam || (w != nullptr && exp_avg_sq != nullptr && step >= 1),
               "corr_linear_wgrad: Adam needs w, exp_avg, exp_avg_sq and step >= 1");
  const uintptr_t al = (uintptr_t)dw | (uintptr_t)w | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq;
  BSIG_REQUIRE((al & 7) == 0, "corr_linear_wgrad: buffers must be 8-byte aligned");
  WgradArgs g;
  g.dy = dy; g.fac = fac; g.ldf = ldf; g.rows = rows;
  g.S = (int)s; g.Q = (int)q; g.F = (int)(s * q + 2);
  g.M = (int)m; g.N = (int)n_out; g.KP = p.KP; g.NP = p.NP; g.npt = p.npt;
  g.num_tiles = p.num_tiles;
  g.dw = dw; g.w = w; g.exp_avg = exp_avg; g.exp_avg_sq = exp_avg_sq;
  const double bc1 = 1.0 - pow((double)beta1, (double)(adam ? step : 1));
  const double bc2 = 1.0 - pow((double)beta2, (double)(adam ? step : 1));
  g.one_minus_b1 = 1.0f - beta1; g.b2 = beta2; g.one_minus_b2 = 1.0f - beta2;
  g.step_size = (float)((double)lr / bc1);
  g.inv_bc2_sqrt = (float)(1.0 / sqrt(bc2));
  g.eps = eps; g.gscale = grad_scale;
  BSIG_CUDA(cudaFuncSetAttribute(corr_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)p.smem));
  corr_wgrad_kernel<<<p.grid, WG_THREADS, p.smem, (cudaStream_t)stream>>>(g, mk_div((uint64_t)q));
  BSIG_LAUNCH_CHECK();
  return 0;
}
