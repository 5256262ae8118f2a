// Latency-oriented fp32 GEMM for the reference-sized layers (minibatch 100,
// hidden 128, a few hundred inputs): every GEMM of an Adam step is ONE launch
// whose critical path is a single round trip to L2.
//
//   * 32x32 output tile, 256 threads x (1x4) accumulators;
//   * the reduction dimension is split over a thread-block CLUSTER (1,1,S<=8)
//     so that each CTA owns <= 64 reduction steps whenever the problem allows:
//     its whole A and B panels (16 loads per thread) are requested at once,
//     together with the epilogue operands (bias / tanh'-source);
//   * partial tiles are summed through distributed shared memory by the
//     cluster's rank-0 CTA in fixed rank order (deterministic), which applies
//     the fused epilogue (bias / tanh / dtanh / sincos) and, for wgrad GEMMs,
//     the fused row sums of A (the bias gradient).
#include <cooperative_groups.h>
#include <stdlib.h>

#include "common.cuh"
#include "gemm.cuh"

namespace cg = cooperative_groups;

namespace bsig {

constexpr int SBM = 32, SBN = 32, SBK = 64;   // SBK = reduction steps per pass
constexpr int APITCH = SBM + 4;               // float4-aligned rows
constexpr int BPITCH = SBN + 4;               // float4-aligned rows

#ifdef BSIG_GS_PROF   // temporary instrumentation: %globaltimer marks of every CTA of the last launch
__device__ unsigned long long gs_prof_buf[256 * 8];
#define GS_MARK(i)                                                                      \
  if (threadIdx.x == 0) {                                                               \
    unsigned long long t_;                                                              \
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                              \
    const int c_ = ((blockIdx.y * gridDim.x + blockIdx.x) * gridDim.z + blockIdx.z);    \
    if (c_ < 256) gs_prof_buf[c_ * 8 + (i)] = t_;                                       \
  }
#else
#define GS_MARK(i)
#endif

// torch.optim.Adam on one element: the arithmetic of adam_kernel (optim.cu), bit for bit.
__device__ __forceinline__ void adam_elem(const GemmArgs& g, float& pp, float gg, float& mm, float& vv) {
  mm = mm + (gg - mm) * g.ad_ob1;
  vv = vv * g.ad_b2 + g.ad_ob2 * gg * gg;
  const float denom = sqrtf(vv) * g.ad_ibc2 + g.ad_eps;
  pp = pp - g.ad_step * (mm / denom);
}
// EPI_ADAM: the parameters this GEMM does not produce gradients for (every other layer: their
// gradients are complete, the weight-gradient GEMM of the first layer is the last kernel of the
// backward pass) are updated by the CTAs that have nothing else left to do.
__device__ __forceinline__ void adam_tail(const GemmArgs& g, int64_t worker, int64_t n_workers) {
  float* p = g.ad_p + g.ad_tail_off;
  float* m = g.ad_m + g.ad_tail_off;
  float* v = g.ad_v + g.ad_tail_off;
  const float* gr = g.ad_g + g.ad_tail_off;
  // four elements per round, every load requested before the first store (one round trip)
  for (int64_t i0 = worker; i0 < g.ad_tail_cnt; i0 += 4 * n_workers) {
    float pp[4], mm[4], vv[4], gg[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t i = i0 + u * n_workers;
      const int64_t ic = i < g.ad_tail_cnt ? i : i0;
      pp[u] = p[ic]; mm[u] = m[ic]; vv[u] = v[ic]; gg[u] = gr[ic];
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t i = i0 + u * n_workers;
      if (i < g.ad_tail_cnt) {
        adam_elem(g, pp[u], gg[u], mm[u], vv[u]);
        p[i] = pp[u]; m[i] = mm[u]; v[i] = vv[u];
      }
    }
  }
}

// Shared memory of a CTA (dynamic, 65 KB): operand tiles of a pass, the four reduction quarters,
// and -- used in rank 0 of a cluster only -- the partial tiles and row sums its peers push in.
struct SmallSmem {
  float As[SBK][APITCH];
  float Bs[SBK][BPITCH];
  float red[4][SBM][BPITCH];
  float inbox[7][SBM][SBN];
  float inbox_rs[7][SBM];
};

// EPI: epilogue (compile-time, keeps tanhf / sincosf out of the other variants);
// GATHER: 0 none, 1 rows of A gathered (a_rows), 2 reduction rows of B gathered.
// All global loads are unconditional (indices clamped into range, out-of-range
// lanes zeroed by a select) so the 16 requests of a pass issue back to back.
template <int EPI, int GATHER>
__global__ void __launch_bounds__(256) gemm_small_kernel(GemmArgs g) {
  extern __shared__ __align__(16) unsigned char gs_smem_raw[];
  SmallSmem& sm = *reinterpret_cast<SmallSmem*>(gs_smem_raw);
  auto& As = sm.As;
  auto& Bs = sm.Bs;
  auto& red = sm.red;
  const int tid = threadIdx.x;
  GS_MARK(0)
  const int tx = tid & 7, ty = tid >> 3;   // epilogue: row ty, cols tx*4..+3
  // inner product: 4x4 outputs per thread (rows 4*my.., cols 4*mx..), the 64 steps of a pass
  // split over four groups of two warps -- two 16-byte shared loads feed 16 FMAs
  const int kg = tid >> 6, mx = tid & 7, my = (tid >> 3) & 7;
  const int i0 = blockIdx.y * SBM, j0 = blockIdx.x * SBN;
  const int S = gridDim.z, rank = blockIdx.z;
  const int r_begin = rank * g.k_per_split;
  const int r_end = min(g.K, r_begin + g.k_per_split);
  const bool want_rsum = (g.rowsum != nullptr) && (blockIdx.x == 0);
  // phase A of the cluster barrier: "my shared memory exists" (waited for before the first push)
  if (S > 1) asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
  const int gi_out = i0 + ty, gj_out = j0 + tx * 4;
  const int gi_c = min(gi_out, g.M - 1);

  // loader mappings: consecutive threads walk the contiguous dimension
  const bool a_r_contig = (g.a_sr == 1);
  const bool b_r_contig = (g.b_sr == 1) && (g.b_sj != 1);
  int a_i[8], a_r[8], b_r[8], b_j[8];
  const float* a_ptr[8];
  const float* b_ptr[8];
  bool a_ok[8], b_ok[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int e = tid + 256 * q;
    // reduction index contiguous in memory: 8 lanes walk one 32-byte sector, 4 sectors a warp,
    // so that the transposing stores (pitch 36) hit 32 different banks
    const int r_c = (tid & 7) | (((tid >> 5) & 7) << 3), i_c = ((tid >> 3) & 3) | (q << 2);
    if (a_r_contig) { a_r[q] = r_c; a_i[q] = i_c; }
    else            { a_i[q] = e & 31; a_r[q] = e >> 5; }
    if (b_r_contig) { b_r[q] = r_c; b_j[q] = i_c; }
    else            { b_j[q] = e & 31; b_r[q] = e >> 5; }
    const int gj = j0 + b_j[q];
    b_ok[q] = gj < g.N;
    b_ptr[q] = g.B + (int64_t)min(gj, g.N - 1) * g.b_sj;
    const int gi = i0 + a_i[q];
    a_ok[q] = gi < g.M;
  }
  // minibatch row indices are written once per training call, before the graph
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int gi = min(i0 + a_i[q], g.M - 1);
    const int64_t row = (GATHER == 1) ? __ldg(g.a_rows + gi) : (int64_t)gi;
    a_ptr[q] = g.A + row * g.a_si;
  }
  // (the same holds for gathered reduction rows: the first pass's indices are fetched here, so
  // that the operand loads after the wait are one round trip, not two)
  int64_t b_row0[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int rr = max(min(r_begin + b_r[q], r_end - 1), 0);
    b_row0[q] = (GATHER == 2) ? __ldg(g.b_rows + rr) : (int64_t)rr;
  }
  // everything above is operand-independent set-up: it overlaps the previous kernel
  pdl_wait_then_release();
  GS_MARK(1)

  // epilogue operands: requested now, consumed after the reduction
  float ep4[4] = {0.f, 0.f, 0.f, 0.f};
  if (EPI == EPI_BIAS || EPI == EPI_BIAS_TANH) {
#pragma unroll
    for (int v = 0; v < 4; ++v) ep4[v] = __ldg(g.bias + min(gj_out + v, g.N - 1));
  } else if (EPI == EPI_MUL_DTANH) {
#pragma unroll
    for (int v = 0; v < 4; ++v)
      ep4[v] = __ldg(g.aux + (int64_t)gi_c * g.ld_aux + min(gj_out + v, g.N - 1));
  }

  float acc4[4][4];
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int v = 0; v < 4; ++v) acc4[u][v] = 0.f;
  float rs = 0.f;
  for (int r0 = r_begin; r0 < r_end; r0 += SBK) {
    float ra[8], rb[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int rr = min(r0 + b_r[q], r_end - 1);
      const int64_t row = (GATHER != 2) ? (int64_t)rr
                          : (r0 == r_begin ? b_row0[q] : (int64_t)__ldg(g.b_rows + rr));
      rb[q] = __ldg(b_ptr[q] + row * g.b_sr);
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int r = min(r0 + a_r[q], r_end - 1);
      ra[q] = __ldg(a_ptr[q] + (int64_t)r * g.a_sr);
    }
    if (r0 != r_begin) __syncthreads();          // previous pass fully consumed
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      As[a_r[q]][a_i[q]] = (a_ok[q] && r0 + a_r[q] < r_end) ? ra[q] : 0.f;
      Bs[b_r[q]][b_j[q]] = (b_ok[q] && r0 + b_r[q] < r_end) ? rb[q] : 0.f;
    }
    __syncthreads();
    GS_MARK(2)
#pragma unroll
    for (int s = 0; s < SBK / 4; ++s) {
      const int kk = kg * (SBK / 4) + s;
      const float4 av = *reinterpret_cast<const float4*>(&As[kk][my * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][mx * 4]);
      const float a4[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        acc4[u][0] = fmaf(a4[u], bv.x, acc4[u][0]);
        acc4[u][1] = fmaf(a4[u], bv.y, acc4[u][1]);
        acc4[u][2] = fmaf(a4[u], bv.z, acc4[u][2]);
        acc4[u][3] = fmaf(a4[u], bv.w, acc4[u][3]);
      }
    }
    if (want_rsum) {       // row sums of A: 8 lanes per row, 8 steps each
#pragma unroll
      for (int j = 0; j < 8; ++j) rs += As[j * 8 + tx][ty];
    }
  }
  // the four reduction quarters meet in shared memory, in a fixed order
#pragma unroll
  for (int u = 0; u < 4; ++u)
    *reinterpret_cast<float4*>(&red[kg][my * 4 + u][mx * 4]) =
        make_float4(acc4[u][0], acc4[u][1], acc4[u][2], acc4[u][3]);
  __syncthreads();
  float acc[4];
  {
    const float4 r0 = *reinterpret_cast<const float4*>(&red[0][ty][tx * 4]);
    const float4 r1 = *reinterpret_cast<const float4*>(&red[1][ty][tx * 4]);
    const float4 r2 = *reinterpret_cast<const float4*>(&red[2][ty][tx * 4]);
    const float4 r3 = *reinterpret_cast<const float4*>(&red[3][ty][tx * 4]);
    acc[0] = (r0.x + r1.x) + (r2.x + r3.x);
    acc[1] = (r0.y + r1.y) + (r2.y + r3.y);
    acc[2] = (r0.z + r1.z) + (r2.z + r3.z);
    acc[3] = (r0.w + r1.w) + (r2.w + r3.w);
  }
  if (want_rsum) {
    rs += __shfl_xor_sync(0xffffffffu, rs, 1);
    rs += __shfl_xor_sync(0xffffffffu, rs, 2);
    rs += __shfl_xor_sync(0xffffffffu, rs, 4);
  }

  GS_MARK(3)
  if (S > 1) {
    // split-K over the cluster: the peers PUSH their partial tiles into rank 0's shared memory
    // (16-byte remote stores, released by the barrier arrival) and leave; rank 0 adds them from
    // its own shared memory in rank order.  No remote read latency, nobody waits for rank 0.
    cg::cluster_group cluster = cg::this_cluster();
    asm volatile("barrier.cluster.wait.aligned;" ::: "memory");            // phase A
    if (rank != 0) {
      SmallSmem* home = cluster.map_shared_rank(&sm, 0);
      *reinterpret_cast<float4*>(&home->inbox[rank - 1][ty][tx * 4]) =
          make_float4(acc[0], acc[1], acc[2], acc[3]);
      if (want_rsum && tx == 0) home->inbox_rs[rank - 1][ty] = rs;
      asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");  // phase B
      if (EPI == EPI_ADAM) {
        const int64_t tile = (int64_t)blockIdx.y * gridDim.x + blockIdx.x;
        adam_tail(g, (tile * (S - 1) + (rank - 1)) * 256 + tid,
                  (int64_t)gridDim.x * gridDim.y * (S - 1) * 256);
      }
      GS_MARK(6)
      return;
    }
    asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");  // phase B
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    GS_MARK(4)
#pragma unroll
    for (int peer = 1; peer < 8; ++peer) {
      if (peer < S) {
        const float4 pv = *reinterpret_cast<const float4*>(&sm.inbox[peer - 1][ty][tx * 4]);
        acc[0] += pv.x; acc[1] += pv.y; acc[2] += pv.z; acc[3] += pv.w;
        if (want_rsum && tx == 0) rs += sm.inbox_rs[peer - 1][ty];
      }
    }
    GS_MARK(5)
  }
  if (EPI == EPI_ADAM && S == 1)
    adam_tail(g, ((int64_t)blockIdx.y * gridDim.x + blockIdx.x) * 256 + tid,
              (int64_t)gridDim.x * gridDim.y * 256);

  if (gi_out >= g.M) return;
  float* c = g.C + (int64_t)gi_out * g.ldc + gj_out;
  if (EPI == EPI_ADAM) {         // all twelve loads before the first store
    const int64_t e0 = (int64_t)gi_out * g.ldc + gj_out;
    float pp[4], mm[4], vv[4];
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const int vc = gj_out + v < g.N ? v : 0;
      pp[v] = c[vc]; mm[v] = g.ad_m[e0 + vc]; vv[v] = g.ad_v[e0 + vc];
    }
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      if (gj_out + v < g.N) {
        adam_elem(g, pp[v], acc[v], mm[v], vv[v]);
        c[v] = pp[v]; g.ad_m[e0 + v] = mm[v]; g.ad_v[e0 + v] = vv[v];
      }
    }
  }
#pragma unroll
  for (int v = 0; v < 4; ++v) {
    if (gj_out + v >= g.N) continue;
    const float a = acc[v];
    if (EPI == EPI_ADAM) continue;
    else if (EPI == EPI_STORE) c[v] = a;
    else if (EPI == EPI_BIAS) c[v] = a + ep4[v];
    else if (EPI == EPI_BIAS_TANH) c[v] = tanhf(a + ep4[v]);
    else if (EPI == EPI_MUL_DTANH) c[v] = a * (1.0f - ep4[v] * ep4[v]);
    else {
      float sn, co;
      sincosf(a, &sn, &co);
      c[v] = g.scale * co;
      c[v + g.N] = g.scale * sn;
    }
  }
  GS_MARK(6)
  if (want_rsum && tx == 0) {
    if (EPI == EPI_ADAM) {       // the bias gradient goes straight into the bias
      const int64_t e = g.ad_b_off + gi_out;
      float pp = g.ad_p[e], mm = g.ad_m[e], vv = g.ad_v[e];
      adam_elem(g, pp, rs, mm, vv);
      g.ad_p[e] = pp; g.ad_m[e] = mm; g.ad_v[e] = vv;
    } else {
      g.rowsum[gi_out] = rs;
    }
  }
}

template <int EPI, int GATHER>
static int launch_small(const GemmArgs& g, dim3 grid, int S, cudaStream_t st) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(256);
  cfg.dynamicSmemBytes = sizeof(SmallSmem);
  cfg.stream = st;
  BSIG_CUDA(cudaFuncSetAttribute(gemm_small_kernel<EPI, GATHER>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)sizeof(SmallSmem)));
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (S > 1) {      // reduction split over a thread-block cluster
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 1;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = (unsigned)S;
    ++na;
  }
  na = add_pdl_attr(attr, na);
  cfg.attrs = attr;
  cfg.numAttrs = na;
  BSIG_CUDA(cudaLaunchKernelEx(&cfg, gemm_small_kernel<EPI, GATHER>, g));
  BSIG_LAUNCH_CHECK();
  return 0;
}

template <int EPI>
static int launch_small_g(const GemmArgs& g, dim3 grid, int S, cudaStream_t st) {
  if (g.a_rows != nullptr && g.b_rows != nullptr) {
    set_error("gemm_small: gathering both operands is not supported");
    return 1;
  }
  if (g.a_rows != nullptr) return launch_small<EPI, 1>(g, grid, S, st);
  if (g.b_rows != nullptr) return launch_small<EPI, 2>(g, grid, S, st);
  return launch_small<EPI, 0>(g, grid, S, st);
}

bool gemm_small_applicable(const GemmArgs& g) {
  // beyond these sizes the tiled / tensor-core engines win
  return (int64_t)g.M * g.N <= 128 * 1024 && g.K <= 8192;
}

int gemm_small(GemmArgs g, cudaStream_t st) {
  BSIG_REQUIRE(g.M >= 1 && g.N >= 1 && g.K >= 1, "gemm: empty problem");
  const int64_t tiles = ceil_div(g.M, SBM) * ceil_div(g.N, SBN);
  static const int max_split = [] {
    const char* e = getenv("BSIG_SMALL_GEMM_MAXSPLIT");   // tuning / experiments only
    const int v = e ? atoi(e) : 8;
    return v < 1 ? 1 : (v > 8 ? 8 : v);
  }();
  // one 64-deep pass per CTA whenever possible (measured: a pass costs ~1.6 us, a
  // cluster of <= 5 CTAs ~0.5 us, larger clusters ~2 us), never more CTAs than 2 waves
  int S = (int)std::min<int64_t>(max_split, std::max<int64_t>(1, (2 * (int64_t)sm_count()) / tiles));
  S = (int)std::min<int64_t>(S, std::max<int64_t>(1, ceil_div(g.K, SBK)));
  int kps = (int)ceil_div(g.K, S);
  kps = (int)(ceil_div(kps, 32) * 32);
  S = (int)ceil_div(g.K, kps);
  g.k_per_split = kps;
  g.partial = nullptr;
  const dim3 grid((unsigned)ceil_div(g.N, SBN), (unsigned)ceil_div(g.M, SBM), (unsigned)S);
  switch (g.epi) {
    case EPI_STORE: return launch_small_g<EPI_STORE>(g, grid, S, st);
    case EPI_BIAS: return launch_small_g<EPI_BIAS>(g, grid, S, st);
    case EPI_BIAS_TANH: return launch_small_g<EPI_BIAS_TANH>(g, grid, S, st);
    case EPI_MUL_DTANH: return launch_small_g<EPI_MUL_DTANH>(g, grid, S, st);
    case EPI_SINCOS: return launch_small_g<EPI_SINCOS>(g, grid, S, st);
    case EPI_ADAM: return launch_small_g<EPI_ADAM>(g, grid, S, st);
  }
  set_error("gemm_small: unknown epilogue %d", g.epi);
  return 1;
}

}  // namespace bsig

#ifdef BSIG_GS_PROF
extern "C" int dbg_gs_prof_read(unsigned long long* out) {
  return (int)cudaMemcpyFromSymbol(out, bsig::gs_prof_buf, sizeof(unsigned long long) * 256 * 8);
}
#endif
