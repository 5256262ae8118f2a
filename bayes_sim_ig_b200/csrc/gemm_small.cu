// Latency-oriented fp32 GEMM for the reference-sized layers (minibatch 100,
// hidden 128, a few hundred inputs): every GEMM of an Adam step is ONE launch.
//
// 32x32 output tile, 128 threads x (2x4) accumulators.  The reduction dimension
// is split over a thread-block CLUSTER (1,1,S<=8): each CTA reduces K/S, the
// partial tiles are summed through distributed shared memory by the cluster's
// rank-0 CTA in fixed rank order (deterministic), which then applies the fused
// epilogue (bias / tanh / dtanh / sincos).  This replaces the two-kernel
// split-K of gemm_simt.cu and turns 4-16 CTAs into 32-128 busy SMs.
// Optional fused row sums of A (the bias gradient of a wgrad GEMM) ride along.
#include <cooperative_groups.h>
#include <stdlib.h>

#include "common.cuh"
#include "gemm.cuh"

namespace cg = cooperative_groups;

namespace bsig {

constexpr int SBM = 32, SBN = 32, SBK = 32;

__device__ __forceinline__ void small_epilogue(const GemmArgs& g, int i, int j, float acc) {
  float* c = g.C + (int64_t)i * g.ldc + j;
  switch (g.epi) {
    case EPI_STORE: *c = acc; break;
    case EPI_BIAS: *c = acc + __ldg(g.bias + j); break;
    case EPI_BIAS_TANH: *c = tanhf(acc + __ldg(g.bias + j)); break;
    case EPI_MUL_DTANH: {
      const float h = __ldg(g.aux + (int64_t)i * g.ld_aux + j);
      *c = acc * (1.0f - h * h);
      break;
    }
    case EPI_SINCOS: {
      float s, co;
      sincosf(acc, &s, &co);
      c[0] = g.scale * co;
      c[g.N] = g.scale * s;
      break;
    }
  }
}

__global__ void __launch_bounds__(128) gemm_small_kernel(GemmArgs g) {
  __shared__ __align__(16) float As[SBK][SBM + 4];
  __shared__ __align__(16) float Bs[SBK][SBN + 4];
  __shared__ float part[SBM][SBN + 1];     // this CTA's partial tile (cluster reduce)
  __shared__ float rsum[SBM];              // partial row sums of A
  const int tid = threadIdx.x;
  const int tx = tid & 7, ty = tid >> 3;   // cols tx*4..+3, rows ty*2, ty*2+1
  const int i0 = blockIdx.y * SBM, j0 = blockIdx.x * SBN;
  const int S = gridDim.z, rank = blockIdx.z;
  const int r_begin = rank * g.k_per_split;
  const int r_end = min(g.K, r_begin + g.k_per_split);
  const bool want_rsum = (g.rowsum != nullptr) && (blockIdx.x == 0);

  const bool a_r_contig = (g.a_sr == 1);
  const bool b_r_contig = (g.b_sr == 1) && (g.b_sj != 1);
  int a_i[8], a_r[8], b_r[8], b_j[8];
  int64_t a_off[8];
  bool a_ok[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int e = tid + 128 * q;
    if (a_r_contig) { a_r[q] = e & 31; a_i[q] = e >> 5; }
    else            { a_i[q] = e & 31; a_r[q] = e >> 5; }
    if (b_r_contig) { b_r[q] = e & 31; b_j[q] = e >> 5; }
    else            { b_j[q] = e & 31; b_r[q] = e >> 5; }
    const int gi = i0 + a_i[q];
    a_ok[q] = gi < g.M;
    const int64_t row = a_ok[q] ? (g.a_rows ? __ldg(g.a_rows + gi) : (int64_t)gi) : 0;
    a_off[q] = row * g.a_si;
  }

  float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
  float rs[2] = {0.f, 0.f};
  float ra[8], rb[8];
  auto fetch = [&](int r0) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int r = r0 + a_r[q];
      ra[q] = (a_ok[q] && r < r_end) ? __ldg(g.A + a_off[q] + (int64_t)r * g.a_sr) : 0.f;
      const int rr = r0 + b_r[q];
      const int gj = j0 + b_j[q];
      float vb = 0.f;
      if (rr < r_end && gj < g.N) {
        const int64_t row = g.b_rows ? __ldg(g.b_rows + rr) : (int64_t)rr;
        vb = __ldg(g.B + row * g.b_sr + (int64_t)gj * g.b_sj);
      }
      rb[q] = vb;
    }
  };
  if (r_begin < r_end) fetch(r_begin);
  for (int r0 = r_begin; r0 < r_end; r0 += SBK) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      As[a_r[q]][a_i[q]] = ra[q];
      Bs[b_r[q]][b_j[q]] = rb[q];
    }
    __syncthreads();
    if (r0 + SBK < r_end) fetch(r0 + SBK);      // prefetch the next slab into registers
#pragma unroll
    for (int kk = 0; kk < SBK; ++kk) {
      const float2 av = *reinterpret_cast<const float2*>(&As[kk][ty * 2]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      acc[0][0] = fmaf(av.x, bv.x, acc[0][0]);
      acc[0][1] = fmaf(av.x, bv.y, acc[0][1]);
      acc[0][2] = fmaf(av.x, bv.z, acc[0][2]);
      acc[0][3] = fmaf(av.x, bv.w, acc[0][3]);
      acc[1][0] = fmaf(av.y, bv.x, acc[1][0]);
      acc[1][1] = fmaf(av.y, bv.y, acc[1][1]);
      acc[1][2] = fmaf(av.y, bv.z, acc[1][2]);
      acc[1][3] = fmaf(av.y, bv.w, acc[1][3]);
      rs[0] += av.x;
      rs[1] += av.y;
    }
    __syncthreads();
  }

  if (S > 1) {
    cg::cluster_group cluster = cg::this_cluster();
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
      for (int v = 0; v < 4; ++v) part[ty * 2 + u][tx * 4 + v] = acc[u][v];
    if (tx == 0) { rsum[ty * 2] = rs[0]; rsum[ty * 2 + 1] = rs[1]; }
    cluster.sync();
    if (rank == 0) {
      for (int peer = 1; peer < S; ++peer) {
        const float* rp = cluster.map_shared_rank(&part[0][0], peer);
#pragma unroll
        for (int u = 0; u < 2; ++u)
#pragma unroll
          for (int v = 0; v < 4; ++v) acc[u][v] += rp[(ty * 2 + u) * (SBN + 1) + tx * 4 + v];
        if (want_rsum && tx == 0) {
          const float* rq = cluster.map_shared_rank(&rsum[0], peer);
          rs[0] += rq[ty * 2];
          rs[1] += rq[ty * 2 + 1];
        }
      }
    }
    cluster.sync();     // peers keep their shared memory alive until rank 0 has read it
    if (rank != 0) return;
  }
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int i = i0 + ty * 2 + u;
    if (i >= g.M) continue;
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const int j = j0 + tx * 4 + v;
      if (j < g.N) small_epilogue(g, i, j, acc[u][v]);
    }
    if (want_rsum && tx == 0) g.rowsum[i] = rs[u];
  }
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
  int S = (int)std::min<int64_t>(max_split, std::max<int64_t>(1, (2 * (int64_t)sm_count()) / tiles));
  S = (int)std::min<int64_t>(S, std::max<int64_t>(1, g.K / SBK));
  int kps = (int)ceil_div(g.K, S);
  kps = (int)(ceil_div(kps, SBK) * SBK);
  S = (int)ceil_div(g.K, kps);
  g.k_per_split = kps;
  g.partial = nullptr;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)ceil_div(g.N, SBN), (unsigned)ceil_div(g.M, SBM), (unsigned)S);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = (unsigned)S;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  BSIG_CUDA(cudaLaunchKernelEx(&cfg, gemm_small_kernel, g));
  BSIG_LAUNCH_CHECK();
  return 0;
}

}  // namespace bsig
