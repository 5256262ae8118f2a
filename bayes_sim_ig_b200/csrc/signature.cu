// Truncated path signature (depth <= 3) of time-augmented rollouts:
// reference bayes_sim_ig/utils/summarizers.py:144-168, whose arithmetic is
// signatory.signature(path, depth) (third-party, absent; restated from the
// published definition -- see oracle/signature_np.py).
//
//   path_t = [t+1 | s_t | a_t]  (C = 1+D+A channels, t = 0..L-1)
//   Sig = exp(d_1) (x) ... (x) exp(d_{L-1}),  d_t = path_t - path_{t-1}
//
// Chen's identity in Horner form, per step with increment d and running levels
// S1,S2,S3 (S1 before the step is path_t - path_0, so it is never stored):
//   S3[i,j,:] += (S2[i,j] + (S1[i] + d[i]/3) * d[j]/2) * d[:]
//   S2[i,j]   += (S1[i] + d[i]/2) * d[j]
// One thread owns one (i,j) pair: S2[i,j] and the C-vector S3[i,j,:] live in
// registers for the whole path; the path itself sits in shared memory.  The
// finished levels are staged in shared memory and leave with coalesced stores.
// ~2C^3 flop per step against 4(L(D+A) + C + C^2 + C^3) bytes per trajectory:
// HBM-bound for small C, FFMA-bound around C = 22.
#include "common.cuh"

namespace bsig {

struct SigArgs {
  const float* states;
  const float* actions;
  float* out;
  int64_t n;
  int64_t s_stride, a_stride;
  int L, D, A, C;
  int tpb;          // trajectories per CTA
  int64_t siglen;
};

__device__ __forceinline__ void load_paths(const SigArgs& p, float* xs, int64_t traj0, int ntraj) {
  // xs[traj][t][c]
  const int per = p.L * p.C;
  for (int e = threadIdx.x; e < ntraj * per; e += blockDim.x) {
    const int tl = e / per, r = e - tl * per;
    const int t = r / p.C, c = r - t * p.C;
    float v;
    if (c == 0) v = (float)(t + 1);
    else if (c <= p.D) v = __ldg(p.states + (traj0 + tl) * p.s_stride + (int64_t)t * p.D + (c - 1));
    else v = __ldg(p.actions + (traj0 + tl) * p.a_stride + (int64_t)t * p.A + (c - 1 - p.D));
    xs[e] = v;
  }
}

template <int CP>
__global__ void __launch_bounds__(512) signature3_kernel(SigArgs p) {
  extern __shared__ float smem[];
  const int C = p.C, L = p.L, CC = C * C;
  float* xs = smem;                                   // [tpb][L][C]
  float* stage = smem + (size_t)p.tpb * L * C;        // [tpb][siglen]
  const int64_t traj0 = (int64_t)blockIdx.x * p.tpb;
  const int ntraj = (int)min((int64_t)p.tpb, p.n - traj0);
  load_paths(p, xs, traj0, ntraj);
  __syncthreads();

  const int tl = threadIdx.x / CC;
  const int ij = threadIdx.x - tl * CC;
  const int i = ij / C, j = ij - i * C;
  if (tl < ntraj) {
    const float* x = xs + (size_t)tl * L * C;
    float s2 = 0.f;
    float s3[CP];
#pragma unroll
    for (int k = 0; k < CP; ++k) s3[k] = 0.f;
    for (int t = 0; t + 1 < L; ++t) {
      const float* x0 = x + t * C;
      const float* x1 = x0 + C;
      const float di = x1[i] - x0[i], dj = x1[j] - x0[j];
      const float s1i = x0[i] - x[i];
      const float t2 = s2 + (s1i + di * (1.0f / 3.0f)) * dj * 0.5f;
#pragma unroll
      for (int k = 0; k < CP; ++k)
        if (k < C) s3[k] = fmaf(t2, x1[k] - x0[k], s3[k]);
      s2 = fmaf(s1i + di * 0.5f, dj, s2);
    }
    float* o = stage + (size_t)tl * p.siglen;
    if (i == 0) o[j] = x[(L - 1) * C + j] - x[j];
    o[C + ij] = s2;
#pragma unroll
    for (int k = 0; k < CP; ++k)
      if (k < C) o[C + CC + ij * C + k] = s3[k];
  }
  __syncthreads();
  float* out = p.out + traj0 * p.siglen;
  const int64_t total = (int64_t)ntraj * p.siglen;
  for (int64_t e = threadIdx.x; e < total; e += blockDim.x) out[e] = stage[e];
}

// depth 2: one CTA per trajectory, threads stride over (i,j); depth 1: differences.
__global__ void __launch_bounds__(256) signature12_kernel(SigArgs p, int depth) {
  extern __shared__ float smem[];
  const int C = p.C, L = p.L, CC = C * C;
  float* x = smem;                                    // [L][C]
  const int64_t traj = blockIdx.x;
  SigArgs q = p;
  load_paths(q, x, traj, 1);
  __syncthreads();
  float* o = p.out + traj * p.siglen;
  for (int c = threadIdx.x; c < C; c += blockDim.x) o[c] = x[(L - 1) * C + c] - x[c];
  if (depth < 2) return;
  for (int ij = threadIdx.x; ij < CC; ij += blockDim.x) {
    const int i = ij / C, j = ij - i * C;
    float s2 = 0.f;
    for (int t = 0; t + 1 < L; ++t) {
      const float* x0 = x + t * C;
      const float* x1 = x0 + C;
      const float di = x1[i] - x0[i], dj = x1[j] - x0[j];
      s2 = fmaf((x0[i] - x[i]) + di * 0.5f, dj, s2);
    }
    o[C + ij] = s2;
  }
}

}  // namespace bsig

using namespace bsig;

extern "C" int bsig_signature_fwd(const float* states, const float* actions, float* out,
                                  int64_t n, int64_t len, int64_t t_states, int64_t t_actions,
                                  int64_t d, int64_t a, int depth, void* stream) {
  BSIG_REQUIRE(n >= 0 && d >= 1 && a >= 0, "signature: bad sizes");
  BSIG_REQUIRE(len >= 2 && t_states >= len && (a == 0 || t_actions >= len),
               "signature: need >= 2 path points and len <= stored steps");
  BSIG_REQUIRE(depth >= 1 && depth <= 3, "signature: depth must be 1, 2 or 3");
  if (n == 0) return 0;
  SigArgs p;
  p.states = states; p.actions = actions; p.out = out; p.n = n;
  p.s_stride = t_states * d; p.a_stride = t_actions * a;
  p.L = (int)len; p.D = (int)d; p.A = (int)a; p.C = (int)(1 + d + a);
  const int64_t C = p.C;
  p.siglen = C + (depth >= 2 ? C * C : 0) + (depth >= 3 ? C * C * C : 0);
  cudaStream_t st = (cudaStream_t)stream;
  if (depth == 3) {
    BSIG_REQUIRE(C <= 22, "signature: depth 3 supports at most 22 channels (got %d)", (int)C);
    const int cc = (int)(C * C);
    int tpb = std::max(1, 256 / cc);
    // keep shared memory under ~100 KB so that two CTAs fit per SM
    const int64_t per_traj = (len * C + p.siglen) * 4;
    tpb = (int)std::max<int64_t>(1, std::min<int64_t>(tpb, (100 * 1024) / per_traj));
    p.tpb = tpb;
    const int threads = (int)(ceil_div((int64_t)tpb * cc, 32) * 32);
    const size_t smem = (size_t)tpb * per_traj;
    const unsigned grid = (unsigned)ceil_div(n, tpb);
#define BSIG_SIG3(CPV)                                                                     \
  do {                                                                                     \
    if (smem > 48 * 1024)                                                                  \
      BSIG_CUDA(cudaFuncSetAttribute(signature3_kernel<CPV>,                               \
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    signature3_kernel<CPV><<<grid, threads, smem, st>>>(p);                                \
  } while (0)
    if (C <= 8) BSIG_SIG3(8);
    else if (C <= 16) BSIG_SIG3(16);
    else BSIG_SIG3(22);
#undef BSIG_SIG3
  } else {
    p.tpb = 1;
    const size_t smem = (size_t)len * C * 4;
    BSIG_REQUIRE(smem <= 200 * 1024, "signature: path too large for shared memory");
    if (smem > 48 * 1024)
      BSIG_CUDA(cudaFuncSetAttribute(signature12_kernel,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    BSIG_REQUIRE(n < (1ll << 31), "signature: too many trajectories for one launch");
    signature12_kernel<<<(unsigned)n, 256, smem, st>>>(p, depth);
  }
  BSIG_LAUNCH_CHECK();
  return 0;
}
