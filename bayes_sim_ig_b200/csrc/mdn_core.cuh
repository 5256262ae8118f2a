// Per-sample core of the mixture-density head epilogue + mixture-of-Gaussians NLL
// (reference bayes_sim_ig/models/mdnn.py:109-119 and 127-178): argument block, lane-group
// reductions and the forward/backward routine shared by the NLL kernels (mdn.cu) and the
// persistent training kernel (train_persistent.cu).
#pragma once

#include <cooperative_groups.h>

#include "common.cuh"

namespace bsig {

constexpr float kLLLimit = 1.0e5f;     // MDNN.LL_LIMIT   mdnn.py:22
constexpr float kMinWeight = 1.0e-5f;  // MDNN.MIN_WEIGHT mdnn.py:23
constexpr float kEpsNoise = 1.0e-5f;   // MDNN.EPS_NOISE  mdnn.py:24
constexpr float kLog2Pi = 1.8378770664093453f;

constexpr int kMaxParts = 1024;  // per-block partial slots
// workspace layout (floats): [0,kMaxParts) exp-sum partials, then loss
// partials, then eps-gradient partials, then 4 counters/scalars.
constexpr int kWsFloats = 3 * kMaxParts + 8;

struct NllArgs {
  // forward inputs (row strides in floats)
  const float* z_pi;   int64_t ld_pi;    // FUSED: logits; else weights [b,K]
  const float* mu;     int64_t ld_mu;
  const float* zd;     int64_t ld_zd;    // FUSED: log-diag; else l_d
  const float* low;    int64_t ld_low;   // nullable
  const float* noise;                    // FUSED only, [b,P,K]
  const float* y;      const int64_t* y_rows;
  const float* grad_scale;               // nullable device scalar (non-fused bwd)
  // outputs
  float* loss;
  float* d_pi;   int64_t ldo_pi;         // FUSED: d logits; else d weights
  float* d_mu;   int64_t ldo_mu;
  float* d_zd;   int64_t ldo_zd;
  float* d_low;  int64_t ldo_low;
  float* ws;
  int* flag;
  int B, P, K, L;
  int nparts_e;                          // number of valid exp-sum partials
};

// Flat index -> (row, column) walker for grid-stride loops over a [rows, width]
// block: one division at start, none per iteration.
struct RowCol {
  int64_t row;
  int col;
  int64_t d_row;
  int d_col;
  int width;
  __device__ __forceinline__ RowCol(int64_t start, int64_t stride, int w) : width(w) {
    row = start / w;
    col = (int)(start - row * w);
    d_row = stride / w;
    d_col = (int)(stride - d_row * w);
  }
  __device__ __forceinline__ void next() {
    row += d_row;
    col += d_col;
    if (col >= width) { col -= width; ++row; }
  }
};

template <int GW>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = GW / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
template <int GW>
__device__ __forceinline__ float group_max(float v) {
#pragma unroll
  for (int o = GW / 2; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// --------------------------------------------------------------------- NLL core
// FUSED: inputs are raw head outputs (softmax / exp+noise applied here, and the
//        gradients are taken back through them to the logits).
// FULL : full covariance (strict-lower block present).
// BWD  : also write gradients.
// Per-sample work shared by the grid-wide and the single-cluster kernels.
// Groups of GW lanes own samples base+gid for base = first, first+stride, ...
// TS = threads per CTA (stride of the per-thread z/v vectors in shared memory).
// Operand load of the per-sample routines: read-only global path, or a plain load when
// the operands were staged in shared memory (nll_stream_kernel).
template <bool SM>
__device__ __forceinline__ float ldf(const float* p) {
  if (SM) return *p;
  return __ldg(p);
}

// Transcendentals / division of the per-sample routines (template flag FAST, which
// defaults to SM).  The shared-memory staged
// streaming kernel (SM) is instruction-issue bound, so it uses the SFU forms
// (ex2.approx / lg2.approx / rcp.approx: <= 2 ulp, i.e. ~2e-7 relative, two orders below
// the 1e-5 parity tolerance); the minibatch kernels keep the full-precision functions.
// In the SM form the per-element finite checks are dropped: a non-finite mu / L_d entry
// always reaches the component's log-density, which is checked.
template <bool SM> __device__ __forceinline__ float xexp(float x) { return SM ? __expf(x) : expf(x); }
template <bool SM> __device__ __forceinline__ float xlog(float x) { return SM ? __logf(x) : logf(x); }
template <bool SM> __device__ __forceinline__ float xdiv(float a, float b) {
  return SM ? __fdividef(a, b) : a / b;
}

// Lanes that own one sample.  GW > 0: aligned groups of GW (power of two) lanes, xor
// butterflies.  GW == 0 (one component per lane, K <= 32): groups of exactly K lanes
// packed floor(32 / K) to a warp -- no padding lanes for K = 10 -- reduced with cyclic
// rotations: window sums of width 1, 2, 4, ... are combined along the binary digits of
// K, then lane 0's total is broadcast so that every lane of the group holds the same bits.
template <int GW>
struct LaneGroup {
  int lane_g, gid;
  bool active;
  __device__ __forceinline__ LaneGroup(int tid, int) : lane_g(tid & (GW - 1)), gid(tid / GW),
                                                       active(true) {}
  __device__ __forceinline__ float sum(float v) const { return group_sum<GW>(v); }
  __device__ __forceinline__ float max(float v) const { return group_max<GW>(v); }
  static __device__ __forceinline__ int per_cta(int threads, int) { return threads / GW; }
};
template <>
struct LaneGroup<0> {
  int lane_g, gid, K, base;
  bool active;
  __device__ __forceinline__ LaneGroup(int tid, int k) : K(k) {
    const int lane = tid & 31, rpw = 32 / k, gi = lane / k;
    lane_g = lane - gi * k;
    active = gi < rpw;
    gid = (tid >> 5) * rpw + gi;
    base = lane - lane_g;
  }
  // lane holding element (lane_g + j) mod K of this group, 0 <= j < K; idle lanes: self
  __device__ __forceinline__ int rot(int j) const {
    int t = lane_g + j;
    if (t >= K) t -= K;
    return active ? base + t : base + lane_g;
  }
  __device__ __forceinline__ float sum(float v) const {
    float w = v, tot = 0.f;
    int off = 0;
    for (int width = 1; width <= K; width <<= 1) {
      if (K & width) {
        tot = off ? tot + __shfl_sync(0xffffffffu, w, rot(off)) : w;
        off += width;
      }
      if (2 * width <= K) w += __shfl_sync(0xffffffffu, w, rot(width));
    }
    return __shfl_sync(0xffffffffu, tot, base);
  }
  __device__ __forceinline__ float max(float v) const {
    for (int width = 1; width < K; width <<= 1)
      v = fmaxf(v, __shfl_sync(0xffffffffu, v, rot(width)));
    return v;
  }
  static __device__ __forceinline__ int per_cta(int threads, int k) {
    return (threads >> 5) * (32 / k);
  }
};

template <int GW, int KPL, bool FUSED, bool FULL, bool BWD, bool SM = false, bool FAST = SM>
__device__ __forceinline__ void nll_samples(const NllArgs& a, const float eps,
                                            const float coef_scale, const int first,
                                            const int stride, const int TS, float* zs, float* vs,
                                            float& loss_acc, float& s_acc, bool& bad) {
  const int tid = threadIdx.x;
  const LaneGroup<GW> grp(tid, a.K);
  const int lane_g = grp.lane_g;
  const int gid = grp.gid;
  const int B = a.B, P = a.P, K = a.K;

  for (int base = first; base < B; base += stride) {
    const int b = base + gid;
    const bool row_ok = (b < B) && grp.active;
    const int64_t bb = row_ok ? b : 0;
    const float* yrow = a.y + (a.y_rows ? __ldg(a.y_rows + bb) : bb) * P;
    const float* mu_r = a.mu + bb * a.ld_mu;
    const float* zd_r = a.zd + bb * a.ld_zd;
    const float* low_r = FULL ? a.low + bb * a.ld_low : nullptr;
    const float* nz_r = FUSED ? a.noise + bb * (int64_t)P * K : nullptr;

    // ---- mixture weights
    float w[KPL], soft[KPL], csum = 1.f;
    if (FUSED) {
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < KPL; ++j) {
        const int k = lane_g + j * GW;
        soft[j] = (row_ok && k < K) ? ldf<SM>(a.z_pi + bb * a.ld_pi + k) : -INFINITY;
        mx = fmaxf(mx, soft[j]);
      }
      mx = grp.max(mx);
      float sm = 0.f;
#pragma unroll
      for (int j = 0; j < KPL; ++j) {
        soft[j] = (soft[j] == -INFINITY) ? 0.f : xexp<FAST>(soft[j] - mx);
        sm += soft[j];
      }
      sm = grp.sum(sm);
      float cs = 0.f;
#pragma unroll
      for (int j = 0; j < KPL; ++j) {
        const int k = lane_g + j * GW;
        soft[j] = xdiv<FAST>(soft[j], sm);
        w[j] = (k < K) ? fminf(fmaxf(soft[j], kMinWeight), 1.0f) : 0.f;  // clamped
        cs += w[j];
      }
      csum = grp.sum(cs);
#pragma unroll
      for (int j = 0; j < KPL; ++j) w[j] = xdiv<FAST>(w[j], csum);
    } else {
#pragma unroll
      for (int j = 0; j < KPL; ++j) {
        const int k = lane_g + j * GW;
        w[j] = (row_ok && k < K) ? ldf<SM>(a.z_pi + bb * a.ld_pi + k) : 0.f;
        soft[j] = 0.f;
      }
    }

    // ---- per-component log density
    float r[KPL], g[KPL];
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < KPL; ++j) {
      const int k = lane_g + j * GW;
      r[j] = -INFINITY;
      g[j] = 0.f;
      if (row_ok && k < K) {
        float quad = 0.f, logdet = 0.f;
        if (!FULL && SM) {
          // staged operands: plain strided walk (column k, stride K) -- the compiler
          // strength-reduces the addresses to one add per operand and element
          const float* pz = zd_r + k;
          const float* pm = mu_r + k;
          const float* pn = nz_r + k;
#pragma unroll 4
          for (int i = 0; i < P; ++i) {
            const float ldv = xexp<FAST>(pz[i * K]) + pn[i * K] * eps;
            const float zi = xdiv<FAST>(yrow[i] - pm[i * K], ldv);
            quad = fmaf(zi, zi, quad);
            logdet += xlog<FAST>(ldv);
          }
        } else if (!FULL) {
          // diagonal covariance: no dependence between rows -> issue the loads of
          // four rows together (memory-level parallelism; the batch is latency-bound)
          for (int i0 = 0; i0 < P; i0 += 4) {
            float zv[4], mv[4], nv[4], yv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int i = min(i0 + u, P - 1);
              zv[u] = ldf<SM>(zd_r + i * K + k);
              mv[u] = ldf<SM>(mu_r + i * K + k);
              nv[u] = FUSED ? ldf<SM>(nz_r + i * K + k) : 0.f;
              yv[u] = ldf<SM>(yrow + i);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              if (i0 + u < P) {
                const float ldv = FUSED ? xexp<FAST>(zv[u]) + nv[u] * eps : zv[u];
                const float zi = xdiv<FAST>(yv[u] - mv[u], ldv);
                quad += zi * zi;
                logdet += xlog<FAST>(ldv);
                if (!SM) bad |= !(finite_f(ldv) && finite_f(mv[u]));
              }
            }
          }
        } else {
          for (int i = 0; i < P; ++i) {
            float ldv = ldf<SM>(zd_r + i * K + k);
            if (FUSED) ldv = xexp<FAST>(ldv) + ldf<SM>(nz_r + i * K + k) * eps;
            const float m = ldf<SM>(mu_r + i * K + k);
            float acc = ldf<SM>(yrow + i) - m;
            const float* lrow = low_r + (int64_t)(i * (i - 1) / 2) * K + k;
            for (int c = 0; c < i; ++c) acc -= ldf<SM>(lrow + c * K) * zs[c * TS];
            const float zi = xdiv<FAST>(acc, ldv);
            zs[i * TS] = zi;
            quad += zi * zi;
            logdet += xlog<FAST>(ldv);
            if (!SM) bad |= !(finite_f(ldv) && finite_f(m));
          }
        }
        const float gj = -0.5f * ((float)P * kLog2Pi + quad) - logdet;
        bad |= !(finite_f(gj) && finite_f(w[j]));
        g[j] = gj;
        const float gc = fminf(fmaxf(gj, -kLLLimit), kLLLimit);
        const float wc = fminf(fmaxf(w[j], kMinWeight), 1.0f);
        r[j] = gc + xlog<FAST>(wc);
        mx = fmaxf(mx, r[j]);
      }
    }
    mx = grp.max(mx);
    float se = 0.f;
#pragma unroll
    for (int j = 0; j < KPL; ++j) se += (r[j] == -INFINITY) ? 0.f : xexp<FAST>(r[j] - mx);
    se = grp.sum(se);
    const float lse = mx + xlog<FAST>(se);
    if (row_ok && lane_g == 0) loss_acc -= lse;

    if (BWD) {
      // ---- gradient wrt weights (and back through clamp/renorm/softmax if FUSED)
      float dw[KPL], coef[KPL];
      float t1 = 0.f;
#pragma unroll
      for (int j = 0; j < KPL; ++j) {
        const int k = lane_g + j * GW;
        dw[j] = 0.f;
        coef[j] = 0.f;
        if (row_ok && k < K) {
          const float rho = xexp<FAST>(r[j] - lse);
          coef[j] = -rho * coef_scale;                // d loss / d r_k
          const float wc = fminf(fmaxf(w[j], kMinWeight), 1.0f);
          const bool in_w = (w[j] >= kMinWeight) && (w[j] <= 1.0f);
          dw[j] = in_w ? xdiv<FAST>(coef[j], wc) : 0.f;
          t1 += dw[j] * w[j];
        }
      }
      if (FUSED) {
        t1 = grp.sum(t1);
        float dp[KPL], t2 = 0.f;
#pragma unroll
        for (int j = 0; j < KPL; ++j) {
          const float dc = xdiv<FAST>(dw[j] - t1, csum);
          const bool in_c = (soft[j] >= kMinWeight) && (soft[j] <= 1.0f);
          dp[j] = in_c ? dc : 0.f;
          t2 += dp[j] * soft[j];
        }
        t2 = grp.sum(t2);
#pragma unroll
        for (int j = 0; j < KPL; ++j) {
          const int k = lane_g + j * GW;
          if (row_ok && k < K) a.d_pi[bb * a.ldo_pi + k] = soft[j] * (dp[j] - t2);
        }
      } else {
#pragma unroll
        for (int j = 0; j < KPL; ++j) {
          const int k = lane_g + j * GW;
          if (row_ok && k < K) a.d_pi[bb * a.ldo_pi + k] = dw[j];
        }
      }

      // ---- gradient wrt mu, l_d, low
#pragma unroll
      for (int j = 0; j < KPL; ++j) {
        const int k = lane_g + j * GW;
        if (!(row_ok && k < K)) continue;
        const bool in_g = (g[j] >= -kLLLimit) && (g[j] <= kLLLimit);
        const float cg = in_g ? coef[j] : 0.f;
        float* dmu_r = a.d_mu + bb * a.ldo_mu;
        float* dzd_r = a.d_zd + bb * a.ldo_zd;
        if (!FULL && SM) {
          const float* pz = zd_r + k;
          const float* pm = mu_r + k;
          const float* pn = nz_r + k;
          float* qm = dmu_r + k;
          float* qz = dzd_r + k;
#pragma unroll 4
          for (int i = 0; i < P; ++i) {
            const float nv = pn[i * K];
            const float e = xexp<FAST>(pz[i * K]);
            const float inv = xdiv<FAST>(1.0f, fmaf(nv, eps, e));
            const float zi = (yrow[i] - pm[i * K]) * inv;
            const float vi = zi * inv;
            const float dld = cg * fmaf(vi, zi, -inv);
            qm[i * K] = cg * vi;
            s_acc = fmaf(dld, nv, s_acc);
            qz[i * K] = e * dld;
          }
        } else if (!FULL) {
          for (int i0 = 0; i0 < P; i0 += 4) {
            float zv[4], mv[4], nv[4], yv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int i = min(i0 + u, P - 1);
              zv[u] = ldf<SM>(zd_r + i * K + k);
              mv[u] = ldf<SM>(mu_r + i * K + k);
              nv[u] = FUSED ? ldf<SM>(nz_r + i * K + k) : 0.f;
              yv[u] = ldf<SM>(yrow + i);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int i = i0 + u;
              if (i < P) {
                const float e = FUSED ? xexp<FAST>(zv[u]) : zv[u];
                const float ldv = FUSED ? e + nv[u] * eps : zv[u];
                const float inv = xdiv<FAST>(1.0f, ldv);
                const float zi = (yv[u] - mv[u]) * inv;
                const float vi = zi * inv;
                const float dld = cg * (vi * zi - inv);
                dmu_r[i * K + k] = cg * vi;
                if (FUSED) {
                  s_acc += dld * nv[u];
                  dzd_r[i * K + k] = e * dld;
                } else {
                  dzd_r[i * K + k] = dld;
                }
              }
            }
          }
        } else {
          float* dlow_r = a.d_low + bb * a.ldo_low;
          if (KPL > 1) {
            // zs holds the LAST component's solve: redo the forward substitution
            for (int i = 0; i < P; ++i) {
              float ldv = ldf<SM>(zd_r + i * K + k);
              if (FUSED) ldv = xexp<FAST>(ldv) + ldf<SM>(nz_r + i * K + k) * eps;
              float acc = ldf<SM>(yrow + i) - ldf<SM>(mu_r + i * K + k);
              const float* lrow = low_r + (int64_t)(i * (i - 1) / 2) * K + k;
              for (int c = 0; c < i; ++c) acc -= ldf<SM>(lrow + c * K) * zs[c * TS];
              zs[i * TS] = xdiv<FAST>(acc, ldv);
            }
          }
          // back substitution v = L^-T z, rows descending
          for (int i = P - 1; i >= 0; --i) {
            const float raw = ldf<SM>(zd_r + i * K + k);
            float e = raw, ldv = raw, nz = 0.f;
            if (FUSED) {
              e = xexp<FAST>(raw);
              nz = ldf<SM>(nz_r + i * K + k);
              ldv = e + nz * eps;
            }
            float acc = zs[i * TS];
            for (int c = i + 1; c < P; ++c)
              acc -= ldf<SM>(low_r + (int64_t)(c * (c - 1) / 2 + i) * K + k) * vs[c * TS];
            const float vi = xdiv<FAST>(acc, ldv);
            vs[i * TS] = vi;
            const float zi = zs[i * TS];
            const float dld = cg * (vi * zi - xdiv<FAST>(1.0f, ldv));
            dmu_r[i * K + k] = cg * vi;
            if (FUSED) {
              s_acc += dld * nz;
              dzd_r[i * K + k] = e * dld;
            } else {
              dzd_r[i * K + k] = dld;
            }
            // d L[i][c] = cg * v_i * z_c for c < i
            float* drow = dlow_r + (int64_t)(i * (i - 1) / 2) * K + k;
            const float cv = cg * vi;
            for (int c = 0; c < i; ++c) drow[c * K] = cv * zs[c * TS];
          }
        }
      }
    }
  }

}

}  // namespace bsig
