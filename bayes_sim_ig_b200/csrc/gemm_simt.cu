// fp32 FFMA GEMM with fused epilogues: the exact-fp32 engine behind
// bsig_linear_* / bsig_rff_features (BSIG_GEMM_SIMT).  It serves the
// latency-bound reference-sized problems (minibatch 100, hidden 128) and is the
// bit-stable comparison point for the tcgen05 engine (gemm_tc.cu).
//
//   C[i,j] = epilogue( sum_r A(i,r) * B(r,j) )
//   A(i,r) = A[ (a_rows ? a_rows[i] : i) * a_si + r * a_sr ]
//   B(r,j) = B[ (b_rows ? b_rows[r] : r) * b_sr + j * b_sj ]
//
// 64x64 output tile, 16-deep K slab, 256 threads x (4x4) accumulators,
// register-prefetched double buffering.  Split-K (gridDim.z) writes fp32
// partials that a second kernel reduces in fixed order (deterministic) before
// the epilogue -- this is what gives the 100 x 105002 x 128 ShadowHand first
// layer enough CTAs to stream its 54 MB weight at HBM speed.
#include "common.cuh"
#include "gemm.cuh"

namespace bsig {

constexpr int BM = 64, BN = 64, BK = 16;

__device__ __forceinline__ void epilogue_store(const GemmArgs& g, int i, int j, float acc) {
  switch (g.epi) {
    case EPI_STORE:
      g.C[(int64_t)i * g.ldc + j] = acc;
      break;
    case EPI_BIAS:
      g.C[(int64_t)i * g.ldc + j] = acc + __ldg(g.bias + j);
      break;
    case EPI_BIAS_TANH:
      g.C[(int64_t)i * g.ldc + j] = tanhf(acc + __ldg(g.bias + j));
      break;
    case EPI_MUL_DTANH: {
      const float h = __ldg(g.aux + (int64_t)i * g.ld_aux + j);
      g.C[(int64_t)i * g.ldc + j] = acc * (1.0f - h * h);
      break;
    }
    case EPI_SINCOS: {
      float s, c;
      sincosf(acc, &s, &c);
      g.C[(int64_t)i * g.ldc + j] = g.scale * c;
      g.C[(int64_t)i * g.ldc + g.N + j] = g.scale * s;
      break;
    }
  }
}

__global__ void __launch_bounds__(256) gemm_simt_kernel(GemmArgs g) {
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int i0 = blockIdx.y * BM, j0 = blockIdx.x * BN;
  const int r_begin = blockIdx.z * g.k_per_split;
  const int r_end = min(g.K, r_begin + g.k_per_split);

  // loader mappings: walk the contiguous dimension with consecutive threads
  const bool a_r_contig = (g.a_sr == 1);
  const bool b_r_contig = (g.b_sr == 1) && (g.b_sj != 1);
  int a_i[4], a_r[4], b_r[4], b_j[4];
  int64_t a_off[4];
  bool a_ok[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int e = tid + 256 * q;
    if (a_r_contig) { a_r[q] = e & 15; a_i[q] = e >> 4; }
    else            { a_i[q] = e & 63; a_r[q] = e >> 6; }
    if (b_r_contig) { b_r[q] = e & 15; b_j[q] = e >> 4; }
    else            { b_j[q] = e & 63; b_r[q] = e >> 6; }
    const int gi = i0 + a_i[q];
    a_ok[q] = gi < g.M;
    const int64_t row = a_ok[q] ? (g.a_rows ? __ldg(g.a_rows + gi) : (int64_t)gi) : 0;
    a_off[q] = row * g.a_si;
  }

  float acc[4][4];
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int v = 0; v < 4; ++v) acc[u][v] = 0.f;

  float ra[4], rb[4];
  auto fetch = [&](int r0) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
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
  auto stash = [&](int buf) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      As[buf][a_r[q]][a_i[q]] = ra[q];
      Bs[buf][b_r[q]][b_j[q]] = rb[q];
    }
  };

  int buf = 0;
  if (r_begin < r_end) {
    fetch(r_begin);
    stash(0);
  }
  __syncthreads();
  for (int r0 = r_begin; r0 < r_end; r0 += BK) {
    const bool more = r0 + BK < r_end;
    if (more) fetch(r0 + BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 av = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
      const float a4[4] = {av.x, av.y, av.z, av.w};
      const float b4[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) acc[u][v] = fmaf(a4[u], b4[v], acc[u][v]);
    }
    if (more) stash(buf ^ 1);
    __syncthreads();
    buf ^= 1;
  }

  if (g.partial != nullptr) {
    float* part = g.partial + (int64_t)blockIdx.z * g.M * g.N;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + ty * 4 + u;
      if (i >= g.M) continue;
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        const int j = j0 + tx * 4 + v;
        if (j < g.N) part[(int64_t)i * g.N + j] = acc[u][v];
      }
    }
    return;
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int i = i0 + ty * 4 + u;
    if (i >= g.M) continue;
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const int j = j0 + tx * 4 + v;
      if (j < g.N) epilogue_store(g, i, j, acc[u][v]);
    }
  }
}

__global__ void __launch_bounds__(256) splitk_reduce_kernel(GemmArgs g, int splits) {
  const int64_t total = (int64_t)g.M * g.N;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (int64_t)gridDim.x * blockDim.x) {
    float acc = 0.f;
    for (int s = 0; s < splits; ++s) acc += __ldcg(g.partial + (int64_t)s * total + e);
    const int i = (int)(e / g.N), j = (int)(e - (int64_t)i * g.N);
    epilogue_store(g, i, j, acc);
  }
}

// column sums (bias gradients): db[j] = sum_i dy[i, j]; deterministic.
__global__ void __launch_bounds__(256)
colsum_kernel(const float* __restrict__ dy, float* __restrict__ db, int M, int N) {
  __shared__ float red[8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + cx;
  float acc = 0.f;
  if (j < N)
    for (int i = ry; i < M; i += 8) acc += __ldg(dy + (int64_t)i * N + j);
  red[ry][cx] = acc;
  __syncthreads();
  if (ry == 0 && j < N) {
    float t = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) t += red[q][cx];
    db[j] = t;
  }
}

__global__ void __launch_bounds__(256)
tanh_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y,
                float* __restrict__ dpre, int64_t count) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count;
       i += (int64_t)gridDim.x * blockDim.x) {
    const float h = __ldg(y + i);
    dpre[i] = __ldg(dy + i) * (1.0f - h * h);
  }
}

int choose_splits(int64_t M, int64_t N, int64_t K) {
  const int64_t tiles = ceil_div(M, BM) * ceil_div(N, BN);
  const int64_t target = 2 * (int64_t)sm_count();
  if (tiles >= target || K <= 4 * BK) return 1;
  int64_t s = std::min<int64_t>(ceil_div(target, tiles), K / (2 * BK));
  return (int)std::max<int64_t>(1, std::min<int64_t>(s, 512));
}

int64_t gemm_simt_ws_bytes(int64_t M, int64_t N, int64_t K) {
  const int s = choose_splits(M, N, K);
  return s > 1 ? (int64_t)s * M * N * (int64_t)sizeof(float) : 0;
}

int gemm_simt(GemmArgs g, void* ws, int64_t ws_bytes, cudaStream_t st) {
  BSIG_REQUIRE(g.M >= 1 && g.N >= 1 && g.K >= 1, "gemm: empty problem");
  int splits = choose_splits(g.M, g.N, g.K);
  if (splits > 1 && (ws == nullptr || ws_bytes < gemm_simt_ws_bytes(g.M, g.N, g.K))) splits = 1;
  int kps = (int)ceil_div(g.K, splits);
  kps = (int)(ceil_div(kps, BK) * BK);
  splits = (int)ceil_div(g.K, kps);
  g.k_per_split = kps;
  g.partial = splits > 1 ? (float*)ws : nullptr;
  dim3 grid((unsigned)ceil_div(g.N, BN), (unsigned)ceil_div(g.M, BM), (unsigned)splits);
  BSIG_REQUIRE(grid.y <= 65535, "gemm: M too large for one launch (%d rows)", g.M);
  gemm_simt_kernel<<<grid, 256, 0, st>>>(g);
  BSIG_LAUNCH_CHECK();
  if (splits > 1) {
    const int64_t total = (int64_t)g.M * g.N;
    const int blocks = (int)std::min<int64_t>(ceil_div(total, 256), (int64_t)sm_count() * 8);
    splitk_reduce_kernel<<<blocks, 256, 0, st>>>(g, splits);
    BSIG_LAUNCH_CHECK();
  }
  return 0;
}

// fixed-order sum of `splits` partial results [splits][M][N] (g.partial) + epilogue; shared
// with the tcgen05 engine's split-K
int gemm_splitk_reduce(const GemmArgs& g, int splits, cudaStream_t st) {
  const int64_t total = (int64_t)g.M * g.N;
  const int blocks = (int)std::min<int64_t>(ceil_div(total, 256), (int64_t)sm_count() * 8);
  splitk_reduce_kernel<<<blocks, 256, 0, st>>>(g, splits);
  BSIG_LAUNCH_CHECK();
  return 0;
}

// Large batches: the rows are split over gridDim.y CTAs per 32-column block (partial sums
// [splits][N] in the workspace), a second launch adds the partials in split order --
// deterministic, and the first pass streams dy at HBM speed instead of on ceil(N/32) CTAs.
__global__ void __launch_bounds__(256)
colsum_partial_kernel(const float* __restrict__ dy, float* __restrict__ part, int M, int N,
                      int rows_per_split) {
  __shared__ float red[8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + cx;
  const int r0 = blockIdx.y * rows_per_split, r1 = min(M, r0 + rows_per_split);
  float acc = 0.f;
  if (j < N)
    for (int i = r0 + ry; i < r1; i += 8) acc += __ldg(dy + (int64_t)i * N + j);
  red[ry][cx] = acc;
  __syncthreads();
  if (ry == 0 && j < N) {
    float t = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) t += red[q][cx];
    part[(int64_t)blockIdx.y * N + j] = t;
  }
}

__global__ void __launch_bounds__(256)
colsum_final_kernel(const float* __restrict__ part, float* __restrict__ db, int splits, int N) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= N) return;
  float t = 0.f;
  for (int s = 0; s < splits; ++s) t += __ldcg(part + (int64_t)s * N + j);
  db[j] = t;
}

int colsum(const float* dy, float* db, int64_t M, int64_t N, void* ws, int64_t ws_bytes,
           cudaStream_t st) {
  const int64_t col_blocks = ceil_div(N, 32);
  int64_t splits = std::min<int64_t>(ceil_div(M, 256), std::max<int64_t>(1, (4 * (int64_t)sm_count()) / col_blocks));
  if (M >= 2048 && splits > 1 && ws != nullptr && ws_bytes >= splits * N * (int64_t)sizeof(float)) {
    const int rps = (int)ceil_div(M, splits);
    splits = ceil_div(M, rps);
    const dim3 grid((unsigned)col_blocks, (unsigned)splits);
    colsum_partial_kernel<<<grid, 256, 0, st>>>(dy, (float*)ws, (int)M, (int)N, rps);
    BSIG_LAUNCH_CHECK();
    colsum_final_kernel<<<(unsigned)ceil_div(N, 256), 256, 0, st>>>((const float*)ws, db, (int)splits, (int)N);
    BSIG_LAUNCH_CHECK();
    return 0;
  }
  colsum_kernel<<<(unsigned)col_blocks, 256, 0, st>>>(dy, db, (int)M, (int)N);
  BSIG_LAUNCH_CHECK();
  return 0;
}

}  // namespace bsig

using namespace bsig;

extern "C" int bsig_tanh_bwd(const float* dy, const float* y, float* dpre, int64_t count,
                             void* stream) {
  if (count <= 0) return 0;
  const int blocks = (int)std::min<int64_t>(ceil_div(count, 256), (int64_t)sm_count() * 8);
  tanh_bwd_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(dy, y, dpre, count);
  BSIG_LAUNCH_CHECK();
  return 0;
}
