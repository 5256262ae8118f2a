// Trajectory summarizers: summary_start / summary_waypts and the
// cross-correlation family (reference bayes_sim_ig/utils/summarizers.py:65-130).
//
// Both are pure streaming kernels bounded by HBM: per trajectory they read
// the leading W (<=10) steps -- a few hundred bytes -- and write F floats.
// The output of a batch is ONE dense [N*F] array, so CTAs own 16-byte aligned
// flat ranges of it (row boundaries fall wherever they fall) and every store
// is a full float4 st.global.cs; the tiny per-trajectory windows are staged
// in shared memory once per CTA.
#include "common.cuh"

namespace bsig {

// ------------------------------------------------------------- summary_start
// flat output index e -> (traj, step, column): out[e] = j < D ? s : a.
template <typename idx_t>
__global__ void __launch_bounds__(256)
summary_start_kernel(const float* __restrict__ states, const float* __restrict__ actions,
                     float* __restrict__ out, idx_t total, uint32_t F, uint32_t DA,
                     uint32_t D, uint32_t A, idx_t s_stride, idx_t a_stride, idx_t s_tstride,
                     idx_t a_tstride) {
  const idx_t n4 = (total + 3) / 4;
  for (idx_t i = (idx_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
       i += (idx_t)gridDim.x * blockDim.x) {
    const idx_t e0 = i * 4;
    idx_t traj = e0 / F;
    uint32_t r = (uint32_t)(e0 - traj * F);
    uint32_t t = r / DA;
    uint32_t j = r - t * DA;
    float v[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      if (e0 + c < total) {
        v[c] = (j < D) ? __ldg(states + traj * s_stride + (idx_t)t * s_tstride + j)
                       : __ldg(actions + traj * a_stride + (idx_t)t * a_tstride + (j - D));
      } else {
        v[c] = 0.f;
      }
      ++j;
      ++r;
      if (j == DA) { j = 0; ++t; }
      if (r == F) { r = 0; t = 0; j = 0; ++traj; }
    }
    if (e0 + 3 < total) {
      st_stream_f4(out + e0, make_float4(v[0], v[1], v[2], v[3]));
    } else {
      for (int c = 0; c < 4 && e0 + c < total; ++c) out[e0 + c] = v[c];
    }
  }
}

// ------------------------------------------------------------ cross-correlation
// One CTA = one chunk (blockIdx.y) of one group of G consecutive trajectories
// (blockIdx.x).  G*F is a multiple of 4 floats so group bases are 16B aligned.
struct CrossArgs {
  const float* states;
  const float* actions;
  float* out;
  int* flag;
  int64_t n;
  int64_t s_stride, a_stride;  // floats between consecutive trajectories in the inputs
  int64_t s_tstride, a_tstride;  // floats between consecutive time steps (D, A when
                                 // trajectory-major; N*D, N*A for time-major buffers)
  int D, A, W, Pn, Qn;         // Pn = W*(D-1), Qn = W*A
  int64_t F;                   // Pn*Qn + 2
  int G;                       // trajectories per group
  int64_t chunk;               // floats per chunk (multiple of 4)
  int use_diff;
};

// Exact n / d for n*ceil(2^40/d) < 2^64 and n*d < 2^40 (checked on the host).
struct FastDiv {
  uint32_t d;
  uint64_t m;
};
__device__ __forceinline__ uint32_t fdiv(uint32_t n, const FastDiv& f) {
  return (uint32_t)(((uint64_t)n * f.m) >> 40);
}

// Every thread owns aligned float4s of the CTA's flat range; (row, p, q) come
// from two multiply-shift divisions, the four products of a float4 share at
// most two state features (s0, s1) and read their action features from a
// wrap-extended copy, so the common case has no per-element branches.
// Finiteness: inputs are checked while staging; with finite inputs a product
// can only be non-finite by overflow, tracked with one max per element.
__global__ void __launch_bounds__(256) crosscorr_kernel(CrossArgs p, FastDiv divF, FastDiv divQ,
                                                        FastDiv divP, FastDiv divD, FastDiv divQx,
                                                        FastDiv divA) {
  extern __shared__ float smem[];
  const int Qx = p.Qn + 3;                  // action features + 3 wrap-around copies
  float* sf = smem;                         // [G][Pn]
  float* af = sf + (size_t)p.G * p.Pn;      // [G][Qn+3]
  float* st = af + (size_t)p.G * Qx;        // [G][2] mean, std

  const int64_t traj0 = (int64_t)blockIdx.x * p.G;
  const int gcnt = (int)min((int64_t)p.G, p.n - traj0);
  const uint32_t group_floats = (uint32_t)gcnt * (uint32_t)p.F;
  const uint32_t c0 = blockIdx.y * (uint32_t)p.chunk;
  if (c0 >= group_floats) return;
  const uint32_t c1 = min(c0 + (uint32_t)p.chunk, group_floats);
  const uint32_t F = (uint32_t)p.F, Qn = (uint32_t)p.Qn;
  const uint32_t PQ = F - 2;

  const int g_lo = (int)(c0 / F);
  const int g_hi = (int)((c1 - 1) / F);
  const int Dm1 = p.D - 1;
  const int nrows = g_hi - g_lo + 1;
  bool bad = false;

  for (uint32_t i = threadIdx.x; i < (uint32_t)(nrows * p.Pn); i += blockDim.x) {
    const uint32_t gl = fdiv(i, divP), e = i - gl * (uint32_t)p.Pn;
    const uint32_t t = fdiv(e, divD), j = e - t * (uint32_t)Dm1;
    const int g = g_lo + (int)gl;
    const float* s = p.states + (traj0 + g) * p.s_stride + t * p.s_tstride + j;
    const float lo = __ldg(s);
    const float v = p.use_diff ? (__ldg(s + 1) - lo) : lo;
    bad |= !finite_f(v);
    sf[g * p.Pn + e] = v;
  }
  for (uint32_t i = threadIdx.x; i < (uint32_t)(nrows * Qx); i += blockDim.x) {
    const uint32_t gl = fdiv(i, divQx), e = i - gl * (uint32_t)Qx;
    const int g = g_lo + (int)gl;
    // action feature q = t*A + k; entries >= Qn wrap around
    const uint32_t q = e >= Qn ? e - Qn : e;
    const uint32_t tq = fdiv(q, divA);
    const float v = __ldg(p.actions + (traj0 + g) * p.a_stride + tq * p.a_tstride +
                          (q - tq * (uint32_t)p.A));
    bad |= !finite_f(v);
    af[g * Qx + e] = v;
  }
  __syncthreads();

  // mean / unbiased std of sf (float64, two passes) for the trajectories whose
  // stat slots fall in this chunk: a thread per trajectory when the feature
  // vector is short, a warp per trajectory when it is long.
  if (p.Pn <= 64) {
    for (int g = g_lo + threadIdx.x; g <= g_hi; g += blockDim.x) {
      const uint32_t slot = (uint32_t)g * F + PQ;
      if (slot + 1 < c0 || slot >= c1) continue;
      const float* v = sf + g * p.Pn;
      double acc = 0.0;
      for (int i = 0; i < p.Pn; ++i) acc += (double)v[i];
      const double mean = acc / (double)p.Pn;
      double sq = 0.0;
      for (int i = 0; i < p.Pn; ++i) {
        const double dlt = (double)v[i] - mean;
        sq += dlt * dlt;
      }
      const float sd = p.Pn < 2 ? 0.f : (float)sqrt(sq / (double)(p.Pn - 1));
      st[g * 2 + 0] = (float)mean;
      st[g * 2 + 1] = sd;
      bad |= !finite_f(sd);
    }
  } else {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
    for (int g = g_lo + warp; g <= g_hi; g += nwarp) {
      const uint32_t slot = (uint32_t)g * F + PQ;
      if (slot + 1 < c0 || slot >= c1) continue;
      const float* v = sf + g * p.Pn;
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
      if (lane == 0) {
        const float sd = p.Pn < 2 ? 0.f : (float)sqrt(sq / (double)(p.Pn - 1));
        st[g * 2 + 0] = (float)mean;
        st[g * 2 + 1] = sd;
        bad |= !finite_f(sd);
      }
    }
  }
  __syncthreads();

  float* out = p.out + traj0 * p.F;
  float vmax = 0.f;
  // ---- main loop: float4s that lie entirely inside one row's product block
  // (the others are skipped here -- no divergent slow path -- and written below)
  const bool fast_ok = Qn >= 4;
  for (uint32_t e0 = c0 + 4u * threadIdx.x; e0 < c1; e0 += 4u * 256u) {
    const uint32_t g = fdiv(e0, divF);
    const uint32_t r = e0 - g * F;
    if (fast_ok && r + 3 < PQ) {
      const uint32_t pi = fdiv(r, divQ);
      const uint32_t qi = r - pi * Qn;
      const float* sfg = sf + g * p.Pn;
      const float* afg = af + g * Qx + qi;
      const float s0 = sfg[pi];
      const float s1 = sfg[min(pi + 1, (uint32_t)p.Pn - 1)];
      const uint32_t nfirst = Qn - qi;       // elements still in feature row pi
      float4 v;
      v.x = s0 * afg[0];
      v.y = (nfirst > 1 ? s0 : s1) * afg[1];
      v.z = (nfirst > 2 ? s0 : s1) * afg[2];
      v.w = (nfirst > 3 ? s0 : s1) * afg[3];
      vmax = fmaxf(fmaxf(vmax, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
      st_stream_f4(out + e0, v);
    }
  }
  // ---- boundary float4s: per row, the (at most three) aligned float4s that start in
  // [row0 + PQ - 3, row0 + F): last products, the two statistics, head of the next row.
  // When the product block cannot use the fast path at all (Qn < 4) every float4 of the
  // chunk is handled here.
  const uint32_t n_items = fast_ok ? (uint32_t)nrows * 3u : (c1 - c0 + 3u) / 4u;
  for (uint32_t it = threadIdx.x; it < n_items; it += 256u) {
    uint32_t e0;
    if (fast_ok) {
      const uint32_t g = (uint32_t)g_lo + it / 3u, m = it % 3u;
      const uint32_t row0 = g * F;
      const uint32_t first = (row0 + (PQ >= 3 ? PQ - 3 : 0) + 3u) & ~3u;   // first aligned e0 >= row0+PQ-3
      e0 = first + 4u * m;
      if (e0 < row0 || e0 >= row0 + F || e0 < c0 || e0 >= c1) continue;
      if (e0 - row0 + 3 < PQ) continue;                                    // a fast float4
    } else {
      e0 = c0 + 4u * it;
    }
    float t[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const uint32_t e = e0 + c;
      t[c] = 0.f;
      if (e < c1) {
        const uint32_t gg = fdiv(e, divF);
        const uint32_t rr = e - gg * F;
        if (rr < PQ) {
          const uint32_t pp = fdiv(rr, divQ);
          t[c] = sf[gg * p.Pn + pp] * af[gg * Qx + (rr - pp * Qn)];
        } else {
          t[c] = st[gg * 2 + (rr - PQ)];
        }
        vmax = fmaxf(vmax, fabsf(t[c]));
      }
    }
    if (e0 + 3 < c1) {
      st_stream_f4(out + e0, make_float4(t[0], t[1], t[2], t[3]));
    } else {
      for (int c = 0; c < 4 && e0 + c < c1; ++c) out[e0 + c] = t[c];
    }
  }
  if (bad || !(vmax <= 3.402823466e38f)) atomicOr(p.flag, 1);
}

static int gcd_i(int a, int b) { return b == 0 ? a : gcd_i(b, a % b); }

}  // namespace bsig

using namespace bsig;

static int summary_start_impl(const float* states, const float* actions, float* out, int64_t n,
                              int64_t d, int64_t a, int64_t max_t, int64_t s_stride,
                              int64_t s_tstride, int64_t a_stride, int64_t a_tstride,
                              int64_t s_extent, int64_t a_extent, cudaStream_t st) {
  if (n == 0) return 0;
  const int64_t DA = d + a, F = max_t * DA, total = n * F;
  BSIG_REQUIRE(F < (1ll << 31), "summary_start: row too wide");
  const int64_t n4 = (total + 3) / 4;
  const int threads = 256;
  const int64_t blocks = std::min<int64_t>(ceil_div(n4, threads), (int64_t)sm_count() * 16);
  if (total < (1ll << 31) && s_extent < (1ll << 31) && a_extent < (1ll << 31)) {
    summary_start_kernel<uint32_t><<<(unsigned)blocks, threads, 0, st>>>(
        states, actions, out, (uint32_t)total, (uint32_t)F, (uint32_t)DA, (uint32_t)d,
        (uint32_t)a, (uint32_t)s_stride, (uint32_t)a_stride, (uint32_t)s_tstride,
        (uint32_t)a_tstride);
  } else {
    summary_start_kernel<uint64_t><<<(unsigned)blocks, threads, 0, st>>>(
        states, actions, out, (uint64_t)total, (uint32_t)F, (uint32_t)DA, (uint32_t)d,
        (uint32_t)a, (uint64_t)s_stride, (uint64_t)a_stride, (uint64_t)s_tstride,
        (uint64_t)a_tstride);
  }
  BSIG_LAUNCH_CHECK();
  return 0;
}

extern "C" int bsig_summary_start(const float* states, const float* actions, float* out,
                                  int64_t n, int64_t t_states, int64_t t_actions,
                                  int64_t d, int64_t a, int64_t max_t, void* stream) {
  BSIG_REQUIRE(n >= 0 && d >= 1 && a >= 0 && max_t >= 1, "summary_start: bad sizes");
  BSIG_REQUIRE(t_states >= max_t && t_actions >= max_t,
               "summary_start: need at least max_t=%lld steps (got %lld states, %lld actions)",
               (long long)max_t, (long long)t_states, (long long)t_actions);
  return summary_start_impl(states, actions, out, n, d, a, max_t, t_states * d, d, t_actions * a,
                            a, n * t_states * d, n * t_actions * a, (cudaStream_t)stream);
}

extern "C" int bsig_summary_start_tm(const float* states, const float* actions, float* out,
                                     int64_t n, int64_t t_states, int64_t t_actions,
                                     int64_t d, int64_t a, int64_t max_t, void* stream) {
  BSIG_REQUIRE(n >= 0 && d >= 1 && a >= 0 && max_t >= 1, "summary_start_tm: bad sizes");
  BSIG_REQUIRE(t_states >= max_t && t_actions >= max_t,
               "summary_start_tm: need at least max_t=%lld steps (got %lld states, %lld actions)",
               (long long)max_t, (long long)t_states, (long long)t_actions);
  return summary_start_impl(states, actions, out, n, d, a, max_t, d, n * d, a, n * a,
                            n * t_states * d, n * t_actions * a, (cudaStream_t)stream);
}

static int crosscorr_impl(const float* states, const float* actions, float* out, int64_t n,
                          int64_t t_states, int64_t t_actions, int64_t d, int64_t a, int64_t w,
                          int use_state_diff, int* nonfinite_flag, bool time_major, void* stream) {
  BSIG_REQUIRE(n >= 0 && d >= 2 && a >= 1 && w >= 1, "crosscorr: need d>=2, a>=1, w>=1");
  BSIG_REQUIRE(t_states >= w && t_actions >= w, "crosscorr: trajectories shorter than w");
  BSIG_REQUIRE(nonfinite_flag != nullptr, "crosscorr: flag pointer required");
  if (n == 0) return 0;
  CrossArgs p;
  p.states = states; p.actions = actions; p.out = out; p.flag = nonfinite_flag;
  p.n = n;
  if (time_major) {
    p.s_stride = d; p.a_stride = a; p.s_tstride = n * d; p.a_tstride = n * a;
  } else {
    p.s_stride = t_states * d; p.a_stride = t_actions * a; p.s_tstride = d; p.a_tstride = a;
  }
  p.D = (int)d; p.A = (int)a; p.W = (int)w;
  p.Pn = (int)(w * (d - 1)); p.Qn = (int)(w * a);
  const int64_t PQ = (int64_t)p.Pn * p.Qn;
  BSIG_REQUIRE(PQ < (1ll << 31), "crosscorr: feature row too wide");
  p.F = PQ + 2;
  p.use_diff = use_state_diff;
  // group size: >= 64 KB of output per CTA where the batch allows it, a multiple
  // that keeps group bases 16B aligned, bounded by the staging budget (64 KB).
  const int galign = 4 / gcd_i((int)(p.F % 4 == 0 ? 4 : p.F % 4), 4);
  int64_t G = ceil_div(16384, p.F);
  const int64_t per_traj_smem = (int64_t)(p.Pn + p.Qn + 3 + 2) * 4;
  const int64_t gmax = std::max<int64_t>(1, (64 * 1024) / per_traj_smem);
  G = std::min(G, gmax);
  G = std::max<int64_t>(galign, ceil_div(G, galign) * galign);
  // do not starve the machine when n is small
  while (G > galign && ceil_div(n, G) < 2 * sm_count()) G -= galign;
  p.G = (int)G;
  const size_t smem = (size_t)G * per_traj_smem;
  BSIG_REQUIRE(smem <= 200 * 1024, "crosscorr: window too large for shared memory");
  const int64_t group_floats = G * p.F;
  // multiply-shift division bounds (FastDiv): n*d < 2^40
  BSIG_REQUIRE(group_floats < (1ll << 31) && group_floats * p.F < (1ll << 40) &&
               PQ * p.Qn < (1ll << 40), "crosscorr: feature row too wide for this kernel");
  p.chunk = group_floats <= 49152 ? ceil_div(group_floats, 4) * 4 : 32768;
  const int64_t nchunk = ceil_div(group_floats, p.chunk);
  const int64_t ngroup = ceil_div(n, G);
  BSIG_REQUIRE(ngroup < (1ll << 31) && nchunk <= 65535, "crosscorr: grid too large");
  dim3 grid((unsigned)ngroup, (unsigned)nchunk);
  FastDiv divF, divQ;
  divF.d = (uint32_t)p.F;
  divF.m = ((1ull << 40) + (uint64_t)p.F - 1) / (uint64_t)p.F;
  auto mk = [](uint64_t d) {
    FastDiv f;
    f.d = (uint32_t)d;
    f.m = ((1ull << 40) + d - 1) / d;
    return f;
  };
  divQ = mk((uint64_t)p.Qn);
  const FastDiv divP = mk((uint64_t)p.Pn), divD = mk((uint64_t)(p.D - 1)),
                divQx = mk((uint64_t)(p.Qn + 3)), divA = mk((uint64_t)p.A);
  // staging indices stay below G*(Pn+Qn+3) <= 16K floats: far inside the FastDiv bound
  if (smem > 48 * 1024)
    BSIG_CUDA(cudaFuncSetAttribute(crosscorr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)smem));
  crosscorr_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(p, divF, divQ, divP, divD, divQx,
                                                              divA);
  BSIG_LAUNCH_CHECK();
  return 0;
}

extern "C" int bsig_summary_crosscorr(const float* states, const float* actions, float* out,
                                      int64_t n, int64_t t_states, int64_t t_actions,
                                      int64_t d, int64_t a, int64_t w, int use_state_diff,
                                      int* nonfinite_flag, void* stream) {
  return crosscorr_impl(states, actions, out, n, t_states, t_actions, d, a, w, use_state_diff,
                        nonfinite_flag, false, stream);
}

extern "C" int bsig_summary_crosscorr_tm(const float* states, const float* actions, float* out,
                                         int64_t n, int64_t t_states, int64_t t_actions,
                                         int64_t d, int64_t a, int64_t w, int use_state_diff,
                                         int* nonfinite_flag, void* stream) {
  return crosscorr_impl(states, actions, out, n, t_states, t_actions, d, a, w, use_state_diff,
                        nonfinite_flag, true, stream);
}
