// Adam over one flat fp32 buffer, minibatch row gather, finiteness flag and
// target normalisation (reference models/mdnn.py:203,221-222,234,245-248 and
// the isfinite asserts at mdnn.py:120-124).  All HBM-bound streaming kernels.
#include "common.cuh"

namespace bsig {

// torch.optim.Adam (single-tensor formulation):
//   m += (g - m) * (1 - b1);  v = v*b2 + (1-b2)*g*g
//   p -= (lr / (1-b1^t)) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
// 28 B/param: reads p,g,m,v (16 B) and writes p,m,v (12 B).
__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
            float* __restrict__ v, int64_t count, float one_minus_b1, float b2,
            float one_minus_b2, float step_size, float inv_bc2_sqrt, float eps, float gscale) {
  const int64_t n4 = count / 4;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  float4* p4 = reinterpret_cast<float4*>(p);
  const float4* g4 = reinterpret_cast<const float4*>(g);
  float4* m4 = reinterpret_cast<float4*>(m);
  float4* v4 = reinterpret_cast<float4*>(v);
  auto upd = [&](float& pp, float gg, float& mm, float& vv) {
    gg *= gscale;
    mm = mm + (gg - mm) * one_minus_b1;
    vv = vv * b2 + one_minus_b2 * gg * gg;
    const float denom = sqrtf(vv) * inv_bc2_sqrt + eps;
    pp = pp - step_size * (mm / denom);
  };
  pdl_wait_then_release();
  for (int64_t i = t0; i < n4; i += stride) {
    float4 pp = p4[i], mm = m4[i], vv = v4[i];
    const float4 gg = g4[i];
    upd(pp.x, gg.x, mm.x, vv.x);
    upd(pp.y, gg.y, mm.y, vv.y);
    upd(pp.z, gg.z, mm.z, vv.z);
    upd(pp.w, gg.w, mm.w, vv.w);
    p4[i] = pp; m4[i] = mm; v4[i] = vv;
  }
  for (int64_t i = n4 * 4 + t0; i < count; i += stride) upd(p[i], g[i], m[i], v[i]);
}

__global__ void __launch_bounds__(256)
gather_rows_kernel(const float* __restrict__ src, int64_t ld, const int64_t* __restrict__ rows,
                   float* __restrict__ out, int64_t n_rows, int64_t width) {
  const int64_t total = n_rows * width;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = e / width, c = e - r * width;
    out[e] = __ldg(src + __ldg(rows + r) * ld + c);
  }
}

__global__ void __launch_bounds__(256)
finite_flag_kernel(const float* __restrict__ x, int64_t count, int* flag) {
  bool bad = false;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count;
       i += (int64_t)gridDim.x * blockDim.x)
    bad |= !finite_f(__ldg(x + i));
  if (bad) atomicOr(flag, 1);
}

__global__ void __launch_bounds__(256)
normalize_rows_kernel(const float* __restrict__ x, const float* __restrict__ lows,
                      const float* __restrict__ highs, float* __restrict__ y, int64_t n_rows,
                      int64_t width) {
  const int64_t total = n_rows * width;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t c = e % width;
    const float lo = __ldg(lows + c);
    y[e] = (__ldg(x + e) - lo) / (__ldg(highs + c) - lo);
  }
}

static int grid_for(int64_t work) {
  return (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(work, 256), (int64_t)sm_count() * 8));
}

}  // namespace bsig

using namespace bsig;

extern "C" int bsig_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                              int64_t count, int64_t step, float lr, float beta1, float beta2,
                              float eps, float grad_scale, void* stream) {
  BSIG_REQUIRE(count >= 0 && step >= 1, "adam_step: bad count/step");
  if (count == 0) return 0;
  BSIG_REQUIRE((((uintptr_t)param | (uintptr_t)grad | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) &
                15) == 0, "adam_step: buffers must be 16-byte aligned");
  // bias corrections in double, as torch computes them on the host
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  const float step_size = (float)((double)lr / bc1);
  const float inv_bc2_sqrt = (float)(1.0 / sqrt(bc2));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid_for(count / 4 + 1));
  cfg.blockDim = dim3(256);
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  cfg.attrs = attr;
  cfg.numAttrs = add_pdl_attr(attr, 0);
  BSIG_CUDA(cudaLaunchKernelEx(&cfg, adam_kernel, param, grad, exp_avg, exp_avg_sq, count,
                               1.0f - beta1, beta2, 1.0f - beta2, step_size, inv_bc2_sqrt, eps,
                               grad_scale));
  BSIG_LAUNCH_CHECK();
  return 0;
}

extern "C" int bsig_gather_rows(const float* src, int64_t ld_src, const int64_t* rows, float* out,
                                int64_t n_rows, int64_t width, void* stream) {
  if (n_rows * width <= 0) return 0;
  gather_rows_kernel<<<grid_for(n_rows * width), 256, 0, (cudaStream_t)stream>>>(
      src, ld_src, rows, out, n_rows, width);
  BSIG_LAUNCH_CHECK();
  return 0;
}

extern "C" int bsig_finite_flag(const float* x, int64_t count, int* flag, void* stream) {
  if (count <= 0) return 0;
  finite_flag_kernel<<<grid_for(count), 256, 0, (cudaStream_t)stream>>>(x, count, flag);
  BSIG_LAUNCH_CHECK();
  return 0;
}

extern "C" int bsig_normalize_rows(const float* x, const float* lows, const float* highs, float* y,
                                   int64_t n_rows, int64_t width, void* stream) {
  if (n_rows * width <= 0) return 0;
  normalize_rows_kernel<<<grid_for(n_rows * width), 256, 0, (cudaStream_t)stream>>>(
      x, lows, highs, y, n_rows, width);
  BSIG_LAUNCH_CHECK();
  return 0;
}
