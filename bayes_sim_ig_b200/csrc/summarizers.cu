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
                     uint32_t D, uint32_t A, idx_t s_stride, idx_t a_stride) {
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
        v[c] = (j < D) ? __ldg(states + traj * s_stride + (idx_t)t * D + j)
                       : __ldg(actions + traj * a_stride + (idx_t)t * A + (j - D));
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
  int64_t s_stride, a_stride;  // floats per trajectory in the inputs
  int D, A, W, Pn, Qn;         // Pn = W*(D-1), Qn = W*A
  int64_t F;                   // Pn*Qn + 2
  int G;                       // trajectories per group
  int64_t chunk;               // floats per chunk (multiple of 4)
  int use_diff;
};

__global__ void __launch_bounds__(256) crosscorr_kernel(CrossArgs p) {
  extern __shared__ float smem[];
  float* sf = smem;                         // [G][Pn]
  float* af = sf + (size_t)p.G * p.Pn;      // [G][Qn]
  float* st = af + (size_t)p.G * p.Qn;      // [G][2] mean, std

  const int64_t traj0 = (int64_t)blockIdx.x * p.G;
  const int gcnt = (int)min((int64_t)p.G, p.n - traj0);
  const int64_t group_floats = (int64_t)gcnt * p.F;
  const int64_t c0 = (int64_t)blockIdx.y * p.chunk;
  if (c0 >= group_floats) return;
  const int64_t c1 = min(c0 + p.chunk, group_floats);

  // trajectories whose data this chunk touches
  const int g_lo = (int)(c0 / p.F);
  const int g_hi = (int)((c1 - 1) / p.F);
  const int Dm1 = p.D - 1;

  for (int g = g_lo; g <= g_hi; ++g) {
    const float* s = p.states + (traj0 + g) * p.s_stride;
    const float* a = p.actions + (traj0 + g) * p.a_stride;
    for (int i = threadIdx.x; i < p.Pn; i += blockDim.x) {
      const int t = i / Dm1, j = i - t * Dm1;
      const float lo = __ldg(s + t * p.D + j);
      sf[g * p.Pn + i] = p.use_diff ? (__ldg(s + t * p.D + j + 1) - lo) : lo;
    }
    for (int i = threadIdx.x; i < p.Qn; i += blockDim.x) {
      // actions of the first W steps are contiguous: [t*A + k]
      af[g * p.Qn + i] = __ldg(a + i);
    }
  }
  __syncthreads();

  // mean / unbiased std of sf for trajectories whose stat slots are in range:
  // one warp per trajectory, float64 accumulation, two passes.
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  const int64_t PQ = (int64_t)p.Pn * p.Qn;
  for (int g = g_lo + warp; g <= g_hi; g += nwarp) {
    const int64_t slot = (int64_t)g * p.F + PQ;
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
      st[g * 2 + 0] = (float)mean;
      st[g * 2 + 1] = p.Pn < 2 ? 0.f : (float)sqrt(sq / (double)(p.Pn - 1));
    }
  }
  __syncthreads();

  float* out = p.out + traj0 * p.F;
  bool bad = false;
  const uint32_t Qn = (uint32_t)p.Qn;
  for (int64_t e0 = c0 + (int64_t)threadIdx.x * 4; e0 < c1; e0 += (int64_t)blockDim.x * 4) {
    int g = (int)(e0 / p.F);
    int64_t r = e0 - (int64_t)g * p.F;
    uint32_t pi = 0, qi = 0;
    if (r < PQ) {
      pi = (uint32_t)r / Qn;  // PQ < 2^31 is checked on the host
      qi = (uint32_t)r - pi * Qn;
    }
    float v[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      if (e0 + c < c1) {
        if (r < PQ) {
          v[c] = sf[g * p.Pn + pi] * af[g * p.Qn + qi];
          if (++qi == Qn) { qi = 0; ++pi; }
        } else {
          v[c] = st[g * 2 + (int)(r - PQ)];
        }
        bad |= !finite_f(v[c]);
        if (++r == p.F) { r = 0; pi = 0; qi = 0; ++g; }
      } else {
        v[c] = 0.f;
      }
    }
    if (e0 + 3 < c1) {
      st_stream_f4(out + e0, make_float4(v[0], v[1], v[2], v[3]));
    } else {
      for (int c = 0; c < 4 && e0 + c < c1; ++c) out[e0 + c] = v[c];
    }
  }
  if (bad) atomicOr(p.flag, 1);
}

static int gcd_i(int a, int b) { return b == 0 ? a : gcd_i(b, a % b); }

}  // namespace bsig

using namespace bsig;

extern "C" int bsig_summary_start(const float* states, const float* actions, float* out,
                                  int64_t n, int64_t t_states, int64_t t_actions,
                                  int64_t d, int64_t a, int64_t max_t, void* stream) {
  BSIG_REQUIRE(n >= 0 && d >= 1 && a >= 0 && max_t >= 1, "summary_start: bad sizes");
  BSIG_REQUIRE(t_states >= max_t && t_actions >= max_t,
               "summary_start: need at least max_t=%lld steps (got %lld states, %lld actions)",
               (long long)max_t, (long long)t_states, (long long)t_actions);
  if (n == 0) return 0;
  const int64_t DA = d + a, F = max_t * DA, total = n * F;
  BSIG_REQUIRE(F < (1ll << 31), "summary_start: row too wide");
  const int64_t n4 = (total + 3) / 4;
  const int threads = 256;
  const int64_t blocks = std::min<int64_t>(ceil_div(n4, threads), (int64_t)sm_count() * 16);
  cudaStream_t st = (cudaStream_t)stream;
  if (total < (1ll << 31) && n * t_states * d < (1ll << 31) && n * t_actions * a < (1ll << 31)) {
    summary_start_kernel<uint32_t><<<(unsigned)blocks, threads, 0, st>>>(
        states, actions, out, (uint32_t)total, (uint32_t)F, (uint32_t)DA, (uint32_t)d,
        (uint32_t)a, (uint32_t)(t_states * d), (uint32_t)(t_actions * a));
  } else {
    summary_start_kernel<uint64_t><<<(unsigned)blocks, threads, 0, st>>>(
        states, actions, out, (uint64_t)total, (uint32_t)F, (uint32_t)DA, (uint32_t)d,
        (uint32_t)a, (uint64_t)(t_states * d), (uint64_t)(t_actions * a));
  }
  BSIG_LAUNCH_CHECK();
  return 0;
}

extern "C" int bsig_summary_crosscorr(const float* states, const float* actions, float* out,
                                      int64_t n, int64_t t_states, int64_t t_actions,
                                      int64_t d, int64_t a, int64_t w, int use_state_diff,
                                      int* nonfinite_flag, void* stream) {
  BSIG_REQUIRE(n >= 0 && d >= 2 && a >= 1 && w >= 1, "crosscorr: need d>=2, a>=1, w>=1");
  BSIG_REQUIRE(t_states >= w && t_actions >= w, "crosscorr: trajectories shorter than w");
  BSIG_REQUIRE(nonfinite_flag != nullptr, "crosscorr: flag pointer required");
  if (n == 0) return 0;
  CrossArgs p;
  p.states = states; p.actions = actions; p.out = out; p.flag = nonfinite_flag;
  p.n = n; p.s_stride = t_states * d; p.a_stride = t_actions * a;
  p.D = (int)d; p.A = (int)a; p.W = (int)w;
  p.Pn = (int)(w * (d - 1)); p.Qn = (int)(w * a);
  const int64_t PQ = (int64_t)p.Pn * p.Qn;
  BSIG_REQUIRE(PQ < (1ll << 31), "crosscorr: feature row too wide");
  p.F = PQ + 2;
  p.use_diff = use_state_diff;
  // group size: ~32 KB of output per group, a multiple that keeps 16B alignment,
  // bounded by the shared-memory staging budget (40 KB).
  const int galign = 4 / gcd_i((int)(p.F % 4 == 0 ? 4 : p.F % 4), 4);
  int64_t G = ceil_div(8192, p.F);
  const int64_t per_traj_smem = (int64_t)(p.Pn + p.Qn + 2) * 4;
  const int64_t gmax = std::max<int64_t>(1, (40 * 1024) / per_traj_smem);
  G = std::min(G, gmax);
  G = std::max<int64_t>(galign, (G / galign) * galign);
  // do not starve the machine when n is small
  while (G > galign && ceil_div(n, G) < 2 * sm_count()) G -= galign;
  p.G = (int)G;
  const size_t smem = (size_t)G * per_traj_smem;
  BSIG_REQUIRE(smem <= 200 * 1024, "crosscorr: window too large for shared memory");
  const int64_t group_floats = G * p.F;
  p.chunk = 8192;
  const int64_t nchunk = ceil_div(group_floats, p.chunk);
  const int64_t ngroup = ceil_div(n, G);
  BSIG_REQUIRE(ngroup < (1ll << 31) && nchunk <= 65535, "crosscorr: grid too large");
  if (smem > 48 * 1024)
    BSIG_CUDA(cudaFuncSetAttribute(crosscorr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)smem));
  dim3 grid((unsigned)ngroup, (unsigned)nchunk);
  crosscorr_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(p);
  BSIG_LAUNCH_CHECK();
  return 0;
}
