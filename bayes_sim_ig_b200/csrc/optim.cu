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

// EXPERIMENTAL (opt-in with BSIG_CHAIN=1, single GPU; written and compiled in round 1, NOT yet
// validated on hardware): the weight-gradient GEMMs of ALL layers of an update and Adam in one
// launch.  After the chain kernel (mdn.cu) has produced dz / dh2 / dh1, the three
// dW_l = dY_l^T X_l are independent reductions over the minibatch (B <= 128: one pass, no
// split), nothing reads the old weights any more, so each CTA forms one 32x32 tile of one
// dW_l (and, in the first column tile, the bias gradient = column sums of dY_l) and applies
// Adam to exactly those parameters in its epilogue.
struct Wgrad3Layer {
  const float* dy; int ld_dy;          // [B, n]
  const float* x; int ld_x;            // [*, k]
  const int64_t* x_rows;               // nullable gather of the rows of x
  int n, k;                            // weight [n, k], bias [n]
  int64_t w_off, b_off;                // offsets of weight / bias in the flat parameter buffer
  int tiles_k, tile0;                  // column tiles of this layer, first linear tile id
};
struct Wgrad3Args {
  Wgrad3Layer layer[3];
  float* p; float* m; float* v;        // flat parameters and Adam moments
  int B;
  float one_minus_b1, b2, one_minus_b2, step_size, inv_bc2_sqrt, eps;
};

__global__ void __launch_bounds__(256) wgrad3_adam_kernel(Wgrad3Args a) {
  __shared__ float As[128][33];                      // dY tile  [batch row][output i]
  __shared__ __align__(16) float Bs[128][36];        // X tile   [batch row][input j]
  const int tid = threadIdx.x, tx = tid & 7, ty = tid >> 3;
  int l = 0;
  if ((int)blockIdx.x >= a.layer[1].tile0) l = 1;
  if ((int)blockIdx.x >= a.layer[2].tile0) l = 2;
  const Wgrad3Layer& L = a.layer[l];
  const int tile = (int)blockIdx.x - L.tile0;
  const int ti = tile / L.tiles_k, tj = tile - ti * L.tiles_k;
  const int i0 = ti * 32, j0 = tj * 32;
  pdl_wait_then_release();
  for (int e = tid; e < a.B * 32; e += 256) {
    const int r = e >> 5, c = e & 31;
    As[r][c] = (i0 + c < L.n) ? __ldg(L.dy + (int64_t)r * L.ld_dy + i0 + c) : 0.f;
    const int64_t xr = L.x_rows ? __ldg(L.x_rows + r) : (int64_t)r;
    Bs[r][c] = (j0 + c < L.k) ? __ldg(L.x + xr * L.ld_x + j0 + c) : 0.f;
  }
  __syncthreads();
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  float rs = 0.f;
  for (int kk = 0; kk < a.B; ++kk) {
    const float av = As[kk][ty];
    const float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
    acc[0] = fmaf(av, bv.x, acc[0]);
    acc[1] = fmaf(av, bv.y, acc[1]);
    acc[2] = fmaf(av, bv.z, acc[2]);
    acc[3] = fmaf(av, bv.w, acc[3]);
    rs += av;
  }
  auto adam = [&](int64_t idx, float g) {
    float mm = a.m[idx], vv = a.v[idx];
    mm = mm + (g - mm) * a.one_minus_b1;
    vv = vv * a.b2 + a.one_minus_b2 * g * g;
    const float denom = sqrtf(vv) * a.inv_bc2_sqrt + a.eps;
    a.p[idx] = a.p[idx] - a.step_size * (mm / denom);
    a.m[idx] = mm;
    a.v[idx] = vv;
  };
  const int i = i0 + ty;
  if (i < L.n) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int j = j0 + tx * 4 + q;
      if (j < L.k) adam(L.w_off + (int64_t)i * L.k + j, acc[q]);
    }
    if (tj == 0 && tx == 0) adam(L.b_off + i, rs);
  }
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

extern "C" int bsig_wgrad3_adam_step(const float* dy0, const float* x0, int64_t ld_x0,
                                     const int64_t* x0_rows, int64_t n0, int64_t k0,
                                     int64_t w_off0, int64_t b_off0,
                                     const float* dy1, const float* x1, int64_t n1, int64_t k1,
                                     int64_t w_off1, int64_t b_off1,
                                     const float* dy2, const float* x2, int64_t n2, int64_t k2,
                                     int64_t w_off2, int64_t b_off2,
                                     float* param, float* exp_avg, float* exp_avg_sq, int64_t b,
                                     int64_t step, float lr, float beta1, float beta2, float eps,
                                     void* stream) {
  BSIG_REQUIRE(b >= 1 && b <= 128 && step >= 1, "wgrad3_adam_step: minibatch must be 1..128 rows");
  Wgrad3Args a;
  const float* dys[3] = {dy0, dy1, dy2};
  const float* xs[3] = {x0, x1, x2};
  const int64_t ns[3] = {n0, n1, n2}, ks[3] = {k0, k1, k2};
  const int64_t lds[3] = {ld_x0, k1, k2};
  const int64_t wo[3] = {w_off0, w_off1, w_off2}, bo[3] = {b_off0, b_off1, b_off2};
  int tiles = 0;
  for (int l = 0; l < 3; ++l) {
    BSIG_REQUIRE(ns[l] >= 1 && ks[l] >= 1, "wgrad3_adam_step: empty layer");
    Wgrad3Layer& L = a.layer[l];
    L.dy = dys[l]; L.ld_dy = (int)ns[l];
    L.x = xs[l]; L.ld_x = (int)lds[l];
    L.x_rows = l == 0 ? x0_rows : nullptr;
    L.n = (int)ns[l]; L.k = (int)ks[l];
    L.w_off = wo[l]; L.b_off = bo[l];
    L.tiles_k = (int)ceil_div(ks[l], 32);
    L.tile0 = tiles;
    tiles += (int)ceil_div(ns[l], 32) * L.tiles_k;
  }
  a.p = param; a.m = exp_avg; a.v = exp_avg_sq; a.B = (int)b;
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  a.one_minus_b1 = 1.0f - beta1; a.b2 = beta2; a.one_minus_b2 = 1.0f - beta2;
  a.step_size = (float)((double)lr / bc1);
  a.inv_bc2_sqrt = (float)(1.0 / sqrt(bc2));
  a.eps = eps;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)tiles);
  cfg.blockDim = dim3(256);
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  cfg.attrs = attr;
  cfg.numAttrs = add_pdl_attr(attr, 0);
  BSIG_CUDA(cudaLaunchKernelEx(&cfg, wgrad3_adam_kernel, a));
  BSIG_LAUNCH_CHECK();
  return 0;
}
