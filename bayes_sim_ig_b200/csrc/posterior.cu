// Posterior algebra on the device: predict_MoGs de-normalisation
// (reference models/mdnn.py:264-288), mixture sampling (utils/pdf.py:61-76,
// 296-300, 465-472) and joint log-density (utils/pdf.py:328-332, 474-491).
#include "common.cuh"

namespace bsig {

// --------------------------------------------------------------------- denorm
__global__ void __launch_bounds__(256)
mog_denorm_kernel(const float* __restrict__ weights, const float* __restrict__ mu, int64_t ld_mu,
                  const float* __restrict__ l_d, int64_t ld_ld, const float* __restrict__ low,
                  int64_t ld_low, const float* __restrict__ lows, const float* __restrict__ highs,
                  float* __restrict__ a_out, float* __restrict__ means_out,
                  float* __restrict__ packed_out, int R, int P, int K) {
  const int L = low ? P * (P - 1) / 2 : 0;
  const int W = P + L;                       // packed width
  const int64_t per_r = (int64_t)K * (1 + P + W);
  const int64_t total = (int64_t)R * per_r;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(e / per_r);
    int64_t q = e - (int64_t)r * per_r;
    if (q < K) {                              // mixture weight
      a_out[(int64_t)r * K + q] = __ldg(weights + (int64_t)r * K + q);
      continue;
    }
    q -= K;
    if (q < (int64_t)K * P) {                 // mean: m*rng + lows
      const int k = (int)(q / P), p = (int)(q - (int64_t)k * P);
      float m = __ldg(mu + (int64_t)r * ld_mu + p * K + k);
      if (lows) m = m * (__ldg(highs + p) - __ldg(lows + p)) + __ldg(lows + p);
      means_out[((int64_t)r * K + k) * P + p] = m;
      continue;
    }
    q -= (int64_t)K * P;
    const int k = (int)(q / W), c = (int)(q - (int64_t)k * W);
    float v;
    int row;
    if (c < P) {
      row = c;
      v = __ldg(l_d + (int64_t)r * ld_ld + c * K + k);
    } else {
      const int l = c - P;
      // row of the l-th strict-lower entry in np.tril_indices order
      row = (int)((1.0f + sqrtf(1.0f + 8.0f * (float)l)) * 0.5f);
      while (row * (row - 1) / 2 > l) --row;
      while ((row + 1) * row / 2 <= l) ++row;
      v = __ldg(low + (int64_t)r * ld_low + l * K + k);
    }
    if (lows) v = (__ldg(highs + row) - __ldg(lows + row)) * v;
    packed_out[((int64_t)r * K + k) * W + c] = v;
  }
}

// ------------------------------------------------------------------- sampling
template <typename A_T>
__global__ void __launch_bounds__(256)
mog_pick_kernel(const A_T* __restrict__ a, const double* __restrict__ u, int32_t* comp_idx,
                int32_t* counts, int64_t n, int K) {
  extern __shared__ int32_t hist[];
  for (int i = threadIdx.x; i < K; i += blockDim.x) hist[i] = 0;
  __syncthreads();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const double ui = u[i];
    A_T c = A_T(0);
    int idx = 0;
    for (int j = 0; j + 1 < K; ++j) {
      c = c + a[j];                          // running sum in a's own dtype
      idx += (ui > (double)c) ? 1 : 0;
    }
    comp_idx[i] = idx;
    atomicAdd(&hist[idx], 1);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < K; i += blockDim.x)
    if (hist[i]) atomicAdd(&counts[i], hist[i]);
}

__global__ void __launch_bounds__(256)
mog_affine_kernel(const double* __restrict__ z, const double* __restrict__ means,
                  const double* __restrict__ cmats, const int32_t* __restrict__ counts,
                  double* __restrict__ samples, int64_t n, int P, int K) {
  extern __shared__ int64_t ends[];          // inclusive prefix of counts
  if (threadIdx.x == 0) {
    int64_t acc = 0;
    for (int k = 0; k < K; ++k) { acc += counts[k]; ends[k] = acc; }
  }
  __syncthreads();
  const int64_t total = n * P;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = e / P;
    const int c = (int)(e - row * P);
    int k = 0;
    while (k < K - 1 && row >= ends[k]) ++k;
    const double* C = cmats + (int64_t)k * P * P;
    const double* zr = z + row * P;
    double acc = 0.0;
    for (int r = 0; r < P; ++r) acc += zr[r] * __ldg(C + r * P + c);
    samples[e] = acc + __ldg(means + k * P + c);
  }
}

// One clipped draw per environment, in draw order (the batched form of
// sim/params_generator.py:115-118: distr.gen(n_samples=1)[0] then np.clip): environment e
// uses its own uniform u[e] and normals z[e,:]; nothing is grouped by component.
template <typename A_T>
__global__ void __launch_bounds__(256)
mog_sample_envs_kernel(const A_T* __restrict__ a, const double* __restrict__ u,
                       const double* __restrict__ z, const double* __restrict__ means,
                       const double* __restrict__ cmats, const double* __restrict__ lows,
                       const double* __restrict__ highs, double* __restrict__ samples,
                       int32_t* __restrict__ comp_idx, int64_t n, int P, int K) {
  const int64_t total = n * P;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = e / P;
    const int c = (int)(e - row * P);
    const double ui = u[row];
    A_T cs = A_T(0);
    int k = 0;
    for (int j = 0; j + 1 < K; ++j) {
      cs = cs + a[j];                        // running sum in a's own dtype (pdf.py:70-73)
      k += (ui > (double)cs) ? 1 : 0;
    }
    if (c == 0 && comp_idx != nullptr) comp_idx[row] = k;
    const double* C = cmats + (int64_t)k * P * P;
    const double* zr = z + row * P;
    double acc = 0.0;
    for (int r = 0; r < P; ++r) acc += zr[r] * __ldg(C + r * P + c);
    acc += __ldg(means + k * P + c);
    if (lows != nullptr) acc = fmin(fmax(acc, __ldg(lows + c)), __ldg(highs + c));
    samples[e] = acc;
  }
}

// Philox4x32-10
__device__ __forceinline__ uint4 philox4x32(uint4 ctr, uint2 key) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}
__device__ __forceinline__ float u01(uint32_t x) { return ((float)(x >> 8) + 0.5f) * (1.0f / 16777216.0f); }

__global__ void __launch_bounds__(256)
mog_sample_philox_kernel(const float* __restrict__ a, const float* __restrict__ means,
                         const float* __restrict__ cmats, int32_t* comp_idx,
                         float* __restrict__ samples, uint64_t seed, int64_t n, int P, int K,
                         const float* __restrict__ lows, const float* __restrict__ highs) {
  const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    uint4 rnd = philox4x32(make_uint4((uint32_t)i, (uint32_t)(i >> 32), 0u, 0u), key);
    const float ui = u01(rnd.x);
    float c = 0.f;
    int k = 0;
    for (int j = 0; j + 1 < K; ++j) { c += __ldg(a + j); k += (ui > c) ? 1 : 0; }
    if (comp_idx) comp_idx[i] = k;
    const float* C = cmats + (int64_t)k * P * P;
    float* out = samples + i * P;
    for (int cidx = 0; cidx < P; ++cidx) out[cidx] = __ldg(means + k * P + cidx);
    // normals two at a time (Box-Muller), accumulated row by row of C
    uint32_t blk = 1;
    for (int r = 0; r < P; r += 2) {
      if (((r >> 1) & 1) == 0) rnd = philox4x32(make_uint4((uint32_t)i, (uint32_t)(i >> 32), blk++, 0u), key);
      const uint32_t x0 = ((r >> 1) & 1) ? rnd.z : rnd.x;
      const uint32_t x1 = ((r >> 1) & 1) ? rnd.w : rnd.y;
      const float rad = sqrtf(-2.0f * logf(u01(x0)));
      float sn, cs;
      sincospif(2.0f * u01(x1), &sn, &cs);
      const float z0 = rad * cs, z1 = rad * sn;
      for (int cidx = 0; cidx < P; ++cidx) {
        float acc = out[cidx] + z0 * __ldg(C + r * P + cidx);
        if (r + 1 < P) acc += z1 * __ldg(C + (r + 1) * P + cidx);
        out[cidx] = acc;
      }
    }
    if (lows != nullptr)
      for (int cidx = 0; cidx < P; ++cidx)
        out[cidx] = fminf(fmaxf(out[cidx], __ldg(lows + cidx)), __ldg(highs + cidx));
  }
}

// ----------------------------------------------------------------- log density
template <typename X_T>
__global__ void __launch_bounds__(128)
mog_logpdf_kernel(const X_T* __restrict__ x, const double* __restrict__ a,
                  const double* __restrict__ log_a, const double* __restrict__ means, const double* __restrict__ precs,
                  const double* __restrict__ logdet, double* __restrict__ out, int64_t m, int P,
                  int K, int log_space) {
  const double log2pi = 1.8378770664093454835606594728112;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m;
       i += (int64_t)gridDim.x * blockDim.x) {
    const X_T* xr = x + i * P;
    double run_max = -INFINITY, run_sum = 0.0, lin = 0.0;
    for (int k = 0; k < K; ++k) {
      const double* mk = means + (int64_t)k * P;
      const double* Pk = precs + (int64_t)k * P * P;
      double q = 0.0;
      for (int j = 0; j < P; ++j) {
        double t = 0.0;
        for (int r = 0; r < P; ++r) t += ((double)xr[r] - __ldg(mk + r)) * __ldg(Pk + r * P + j);
        q += t * ((double)xr[j] - __ldg(mk + j));
      }
      const double lp = 0.5 * (-q + __ldg(logdet + k) - (double)P * log2pi);
      if (log_space) {
        const double t = lp + __ldg(log_a + k);
        // a zero-weight component has log_a = -inf: it contributes nothing (scipy's
        // logsumexp, pdf.py:489); without the guard exp(-inf - (-inf)) = NaN would
        // poison the running sum when such a component comes first
        if (t == -INFINITY) continue;
        if (t > run_max) {
          run_sum = run_sum * exp(run_max - t) + 1.0;
          run_max = t;
        } else {
          run_sum += exp(t - run_max);
        }
      } else {
        lin += __ldg(a + k) * exp(lp);
      }
    }
    out[i] = log_space ? run_max + log(run_sum) : lin;
  }
}

// ------------------------------------------------- pairwise marginal densities on grids
// All 2-D marginals a posterior plot needs (utils/plot.py:38-44, 131-149: for every pair of
// parameters, MoG.eval(grid, ii=pair, log=False) on a 100 x 100 np.mgrid) in ONE launch.
// params [NP][K][6] = {m0, m1, p00, p01, p11, logdetP} of each component's 2-D marginal
// (precision of the jittered 2 x 2 covariance block, built on the host); lims [NP][4] =
// {xmin, xmax, ymin, ymax}; grid point (i, j) of pair q is (xmin + i dx, ymin + j dy) with
// dx = (xmax - xmin) / (nbins - 1) -- np.mgrid[xmin:xmax:nbins*1j] -- and lands at
// out[q][i][j].
__global__ void __launch_bounds__(256)
mog_marginal_grid_kernel(const double* __restrict__ a, const double* __restrict__ log_a,
                         const double* __restrict__ params, const double* __restrict__ lims,
                         double* __restrict__ out, int64_t n_pairs, int K, int nbins,
                         int log_space) {
  const double log2pi = 1.8378770664093454835606594728112;
  const int64_t per = (int64_t)nbins * nbins;
  const int64_t total = n_pairs * per;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t q = e / per;
    const int r = (int)(e - q * per);
    const int i = r / nbins, j = r - i * nbins;
    const double* lm = lims + q * 4;
    const double den = (double)(nbins > 1 ? nbins - 1 : 1);
    const double x0 = __ldg(lm + 0) + (double)i * ((__ldg(lm + 1) - __ldg(lm + 0)) / den);
    const double x1 = __ldg(lm + 2) + (double)j * ((__ldg(lm + 3) - __ldg(lm + 2)) / den);
    double run_max = -INFINITY, run_sum = 0.0, lin = 0.0;
    for (int k = 0; k < K; ++k) {
      const double* pk = params + (q * K + k) * 6;
      const double d0 = x0 - __ldg(pk + 0), d1 = x1 - __ldg(pk + 1);
      const double quad = d0 * d0 * __ldg(pk + 2) + 2.0 * d0 * d1 * __ldg(pk + 3) + d1 * d1 * __ldg(pk + 4);
      const double lp = 0.5 * (-quad + __ldg(pk + 5) - 2.0 * log2pi);
      if (log_space) {
        const double t = lp + __ldg(log_a + k);
        if (t == -INFINITY) continue;
        if (t > run_max) {
          run_sum = run_sum * exp(run_max - t) + 1.0;
          run_max = t;
        } else {
          run_sum += exp(t - run_max);
        }
      } else {
        lin += __ldg(a + k) * exp(lp);
      }
    }
    out[e] = log_space ? run_max + log(run_sum) : lin;
  }
}

static int grid_for(int64_t work, int threads) {
  return (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(work, threads), (int64_t)sm_count() * 8));
}

}  // namespace bsig

using namespace bsig;

extern "C" int bsig_mog_denorm(const float* weights, const float* mu, int64_t ld_mu,
                               const float* l_d, int64_t ld_ld, const float* low, int64_t ld_low,
                               const float* lows, const float* highs, float* a_out,
                               float* means_out, float* packed_out, int64_t r, int64_t p,
                               int64_t k, void* stream) {
  BSIG_REQUIRE(r >= 1 && p >= 1 && k >= 1, "mog_denorm: bad sizes");
  BSIG_REQUIRE((lows == nullptr) == (highs == nullptr), "mog_denorm: lows/highs must both be set");
  const int64_t L = low ? p * (p - 1) / 2 : 0;
  const int64_t total = r * k * (1 + p + p + L);
  mog_denorm_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
      weights, mu, ld_mu, l_d, ld_ld, low, ld_low, lows, highs, a_out, means_out, packed_out,
      (int)r, (int)p, (int)k);
  BSIG_LAUNCH_CHECK();
  return 0;
}

extern "C" int bsig_mog_sample(const void* a, int a_is_f32, const double* u, const double* z,
                               const double* means, const double* cmats, int32_t* comp_idx,
                               int32_t* counts, double* samples, int64_t n, int64_t p, int64_t k,
                               void* stream) {
  BSIG_REQUIRE(n >= 0 && p >= 1 && k >= 1 && k <= 8192, "mog_sample: bad sizes");
  if (n == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = grid_for(n, 256);
  if (a_is_f32)
    mog_pick_kernel<float><<<grid, 256, k * sizeof(int32_t), st>>>((const float*)a, u, comp_idx,
                                                                   counts, n, (int)k);
  else
    mog_pick_kernel<double><<<grid, 256, k * sizeof(int32_t), st>>>((const double*)a, u, comp_idx,
                                                                    counts, n, (int)k);
  BSIG_LAUNCH_CHECK();
  mog_affine_kernel<<<grid_for(n * p, 256), 256, k * sizeof(int64_t), st>>>(
      z, means, cmats, counts, samples, n, (int)p, (int)k);
  BSIG_LAUNCH_CHECK();
  return 0;
}

extern "C" int bsig_mog_sample_philox(const float* a, const float* means, const float* cmats,
                                      int32_t* comp_idx, float* samples, uint64_t seed, int64_t n,
                                      int64_t p, int64_t k, void* stream) {
  BSIG_REQUIRE(n >= 0 && p >= 1 && k >= 1, "mog_sample_philox: bad sizes");
  if (n == 0) return 0;
  mog_sample_philox_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(
      a, means, cmats, comp_idx, samples, seed, n, (int)p, (int)k, nullptr, nullptr);
  BSIG_LAUNCH_CHECK();
  return 0;
}

extern "C" int bsig_mog_logpdf(const void* x, int x_is_f32, const double* a, const double* log_a,
                               const double* means, const double* precs, const double* logdet_p, double* out,
                               int64_t m, int64_t p, int64_t k, int log_space, void* stream) {
  BSIG_REQUIRE(m >= 0 && p >= 1 && k >= 1, "mog_logpdf: bad sizes");
  if (m == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = grid_for(m, 128);
  if (x_is_f32)
    mog_logpdf_kernel<float><<<grid, 128, 0, st>>>((const float*)x, a, log_a, means, precs, logdet_p, out,
                                                   m, (int)p, (int)k, log_space);
  else
    mog_logpdf_kernel<double><<<grid, 128, 0, st>>>((const double*)x, a, log_a, means, precs, logdet_p,
                                                    out, m, (int)p, (int)k, log_space);
  BSIG_LAUNCH_CHECK();
  return 0;
}

extern "C" int bsig_mog_sample_envs(const void* a, int a_is_f32, const double* u, const double* z,
                                    const double* means, const double* cmats, const double* lows,
                                    const double* highs, double* samples, int32_t* comp_idx,
                                    int64_t n, int64_t p, int64_t k, void* stream) {
  BSIG_REQUIRE(n >= 0 && p >= 1 && k >= 1, "mog_sample_envs: bad sizes");
  BSIG_REQUIRE((lows == nullptr) == (highs == nullptr), "mog_sample_envs: lows/highs must both be set");
  if (n == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = grid_for(n * p, 256);
  if (a_is_f32)
    mog_sample_envs_kernel<float><<<grid, 256, 0, st>>>((const float*)a, u, z, means, cmats, lows,
                                                        highs, samples, comp_idx, n, (int)p, (int)k);
  else
    mog_sample_envs_kernel<double><<<grid, 256, 0, st>>>((const double*)a, u, z, means, cmats, lows,
                                                         highs, samples, comp_idx, n, (int)p, (int)k);
  BSIG_LAUNCH_CHECK();
  return 0;
}

extern "C" int bsig_mog_sample_envs_philox(const float* a, const float* means, const float* cmats,
                                           const float* lows, const float* highs, int32_t* comp_idx,
                                           float* samples, uint64_t seed, int64_t n, int64_t p,
                                           int64_t k, void* stream) {
  BSIG_REQUIRE(n >= 0 && p >= 1 && k >= 1, "mog_sample_envs_philox: bad sizes");
  BSIG_REQUIRE((lows == nullptr) == (highs == nullptr),
               "mog_sample_envs_philox: lows/highs must both be set");
  if (n == 0) return 0;
  mog_sample_philox_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(
      a, means, cmats, comp_idx, samples, seed, n, (int)p, (int)k, lows, highs);
  BSIG_LAUNCH_CHECK();
  return 0;
}

extern "C" int bsig_mog_marginal_grid(const double* a, const double* log_a, const double* params,
                                      const double* lims, double* out, int64_t n_pairs, int64_t k,
                                      int64_t nbins, int log_space, void* stream) {
  BSIG_REQUIRE(n_pairs >= 0 && k >= 1 && nbins >= 1, "mog_marginal_grid: bad sizes");
  if (n_pairs == 0) return 0;
  mog_marginal_grid_kernel<<<grid_for(n_pairs * nbins * nbins, 256), 256, 0, (cudaStream_t)stream>>>(
      a, log_a, params, lims, out, n_pairs, (int)k, (int)nbins, log_space);
  BSIG_LAUNCH_CHECK();
  return 0;
}
