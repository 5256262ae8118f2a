// Persistent training kernel: ALL Adam updates of an MDNN / MDRFF `run_training` call
// (reference bayes_sim_ig/models/mdnn.py:217-234: minibatch gather -> forward -> mixture NLL
// -> backward -> Adam) inside ONE launch of ONE thread-block cluster.
//
// Why: at the reference's sizes (minibatch 100, 128-wide layers, 90 k parameters) an update is
// ~54 MFLOP of dependent small GEMMs; as separate launches it is bound by launch + L2 round-trip
// latency (10 launches, 47 us per update in round 1).  Here the 16 CTAs of a (non-portable)
// cluster keep the model RESIDENT in shared memory for the whole call, partitioned by OUTPUT
// COLUMN of every layer (weight-stationary):
//
//   CTA c owns columns [c*cpc, (c+1)*cpc) of each dense layer (its rows of W, its bias entries
//   and, in global memory, its slice of the Adam moments), so
//     forward   h_l[:, own] = act(in_l . W_own^T + b_own)           needs the full input in_l
//     wgrad     dW_own = dY_l[:, own]^T . in_l ; Adam on W_own       is local
//     dgrad     partial d in_l = dY_l[:, own] . W_own                summed over the 16 CTAs
//   and the only cross-CTA traffic is activations: an all-gather of each layer's output
//   (forward) and a reduce-scatter of each dgrad partial (backward), both through L2-resident
//   global scratch ordered by hardware cluster barriers (barrier.cluster release/acquire;
//   measured on B200: ~450 cycles per barrier, 60-95 B/clk per SM from L2 -- DSMEM pushes were
//   measured 2-4x slower for these sizes, profiles/r2/microbench_design.txt).
//   The mixture NLL (softmax/clamp/renorm, exp + eps-noise, Cholesky log-density, logsumexp,
//   and their backward: csrc/mdn_core.cuh) is row-parallel: CTA c takes rows [c*rpc, ...) of
//   the gathered head output; the three batch-wide sums travel with the same barriers.
//
// Shared memory (per CTA, <= 227 KB): weight slices | own-column copies of h_l (later dY_l) |
// R_H: one gathered hidden activation [B][pad] | R_X: the gathered minibatch x [B][pad]
// (bulk async copies, one per row) which doubles as reduction scratch while x is dead.
// All arithmetic is fp32 FFMA with fixed-order reductions: results are deterministic and agree
// with the launch-per-GEMM path to summation order.
//
// Code footprint is a first-order design constraint: an update executes every phase once, so
// the per-update instruction stream must stay resident in the SM's instruction cache (the
// first version -- three tile widths, everything unrolled, 150 KB of SASS -- spent its time
// in instruction-fetch misses: profiles/r2/).  Hence ONE column-tile width (8), one instance
// of each tile routine, rolled loops, run-time mixture-group width.
#include <cooperative_groups.h>

#include <algorithm>

#include "async_copy.cuh"
#include "common.cuh"
#include "mdn_core.cuh"

namespace cg = cooperative_groups;

namespace bsig {

constexpr int kTpThreads = 512;
constexpr int kTpWarps = kTpThreads / 32;
constexpr int kTpNC = 16;          // CTAs per cluster (non-portable size)
constexpr int kTpMaxLayers = 3;
constexpr int kTpT = 8;            // column tile width (forward and weight gradient)

struct TpLayer {
  int K, N;            // input / output width of the dense layer
  int cpc;             // output columns per CTA = ceil(N / NC)
  int ldy;             // cpc rounded up to the column tile (multiple of 8)
  int ldw;             // K rounded up to 4: pitch of the weight slice in shared memory
  int lda;             // pitch of this layer's INPUT activation in shared memory
  long long w_off, b_off;   // float offsets of W [N,K] / b [N] in the flat parameter buffer
  int sw, sb;          // shared-memory offsets (floats): weight slice [ldy][ldw], bias [ldy]
  int own;             // shared-memory offset: own-column outputs, later gradients [B][ldy]
  long long g_act;     // global scratch: gathered output [B][N] (hidden) or z [B][NHp] (head)
  long long g_part;    // global scratch: dgrad partials of this layer's input [NC][B][K]
};

struct TpArgs {
  const float* x; long long ldx;       // training set [n_train][ldx], ldx % 4 == 0, pad zero
  const float* y;                      // normalised targets [n_train][P]
  const long long* idx;                // minibatch rows [n_updates][B]
  const float* noise;                  // eps-noise uniforms [n_updates][B][P][K]
  float* params;                       // flat parameters
  float* gx;                           // global exchange scratch
  float* loss; const int* loss_slot;   // loss_buf, per-update slot (-1: none)
  const float* adam_coef;              // [n_updates][2] = step_size, 1/sqrt(bias_correction2)
  int* flag;
  int B, F, P, Kc, L, NH, NHp, gw;     // gw: power of two >= Kc (lanes of a mixture group)
  int n_layers;
  TpLayer layer[kTpMaxLayers];
  long long g_dz, g_sum, g_mv;         // global scratch: dz [B][NHp]; sums [3][NC]; Adam moments
  int wslice;                          // floats of one CTA's weight + bias slices (= moments)
  int rh, rx;                          // shared-memory offsets (floats) of R_H, R_X
  int rh_floats, rx_floats;
  int ldxs;                            // pitch of x in R_X
  int rpc;                             // NLL rows per CTA
  int step0, step1;
  float one_minus_b1, b2, one_minus_b2, eps;
  long long* prof;                     // nullable: per-phase cycle counters of CTA 0
};

__device__ __forceinline__ void cluster_arrive() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void cluster_wait() {
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__device__ __forceinline__ float4 ldcg4(const float* p) {
  return __ldcg(reinterpret_cast<const float4*>(p));
}
__device__ __forceinline__ void fma4(float (&acc)[4], float d, const float4& a) {
  acc[0] = fmaf(d, a.x, acc[0]);
  acc[1] = fmaf(d, a.y, acc[1]);
  acc[2] = fmaf(d, a.z, acc[2]);
  acc[3] = fmaf(d, a.w, acc[3]);
}
__device__ __forceinline__ float dot4(const float4& a, const float4& w, float acc) {
  acc = fmaf(a.x, w.x, acc);
  acc = fmaf(a.y, w.y, acc);
  acc = fmaf(a.z, w.z, acc);
  return fmaf(a.w, w.w, acc);
}

struct AdamC {
  float one_minus_b1, b2, one_minus_b2, eps, step_size, inv_bc2;
};
__device__ __forceinline__ void adam1(float& w, float g, float& m, float& v, const AdamC& c) {
  m = m + (g - m) * c.one_minus_b1;
  v = v * c.b2 + c.one_minus_b2 * g * g;
  const float denom = sqrtf(v) * c.inv_bc2 + c.eps;
  w = w - c.step_size * __fdividef(m, denom);
}

// ------------------------------------------------------------------------------ forward
// One 8-column tile of a forward layer: out[r][n] = act(sum_k A[r][k] W[n][k] + bias[n]).
// Main loop: lanes <-> rows (r = lane + 32 i, i < 4), warps <-> k-quads (q = warp, warp+16,
// ...; W reads are warp-wide broadcasts).  The 16 per-warp partial tiles meet in `scratch`
// ([16][8][B], row fastest: conflict-free) and the epilogue -- thread <-> (row, n mod 4) --
// adds them in slice order, applies bias / tanh and writes the own-column copy and the
// gathered global copy.  Returns this thread's sum of exp(out) over the columns inside
// [e_lo, e_hi) (global column index; head layer: the log-diagonal block).
__device__ __forceinline__ float fwd_tile(const float* __restrict__ A, int lda, int B,
                                          const float* __restrict__ W, int ldw,
                                          const float* __restrict__ bias, int nvalid, bool tanh_act,
                                          float* __restrict__ own, int ldy,
                                          float* __restrict__ gout, int ldg, int gcol0,
                                          int e_lo, int e_hi, float* __restrict__ scratch) {
  constexpr int TN = kTpT;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float acc[4][TN];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int n = 0; n < TN; ++n) acc[i][n] = 0.f;
  const float* a0 = A + (size_t)min(lane, B - 1) * lda;
  const float* a1 = A + (size_t)min(lane + 32, B - 1) * lda;
  const float* a2 = A + (size_t)min(lane + 64, B - 1) * lda;
  const float* a3 = A + (size_t)min(lane + 96, B - 1) * lda;
  const int nq = ldw >> 2;
#pragma unroll 1
  for (int q = warp; q < nq; q += kTpWarps) {
    float4 a[4];
    a[0] = *reinterpret_cast<const float4*>(a0 + 4 * q);
    a[1] = *reinterpret_cast<const float4*>(a1 + 4 * q);
    a[2] = *reinterpret_cast<const float4*>(a2 + 4 * q);
    a[3] = *reinterpret_cast<const float4*>(a3 + 4 * q);
#pragma unroll
    for (int n = 0; n < TN; ++n) {
      const float4 w = *reinterpret_cast<const float4*>(W + (size_t)n * ldw + 4 * q);
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i][n] = dot4(a[i], w, acc[i][n]);
    }
  }
  float* s = scratch + (size_t)warp * TN * B + lane;
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (lane + 32 * i < B) {
#pragma unroll
      for (int n = 0; n < TN; ++n) s[n * B + 32 * i] = acc[i][n];
    }
  __syncthreads();
  float esum = 0.f;
  const int r = tid & 127;
  if (r < B) {
#pragma unroll 1
    for (int n = tid >> 7; n < nvalid; n += 4) {
      const float* p = scratch + (size_t)n * B + r;
      float v = 0.f;
#pragma unroll
      for (int q = 0; q < kTpWarps; ++q) v += p[(size_t)q * TN * B];
      v += bias[n];
      if (tanh_act) v = tanhf(v);
      own[r * ldy + n] = v;
      __stcg(gout + (size_t)r * ldg + n, v);
      const int col = gcol0 + n;
      if (col >= e_lo && col < e_hi) esum += expf(v);
    }
  }
  __syncthreads();
  return esum;
}

// ------------------------------------------------------------------------------- wgrad
// One 8-column tile of dW_own[c][j] = sum_b dY[b][c] A[b][j] and Adam on those weights.
// Main loop: threads <-> (j-quad, batch split); partial tiles to scratch [BS][8][ldw]; then
// thread <-> float4 of the tile: sum the splits, Adam (moments live in global scratch in the
// SAME [c][ldw] layout as the shared-memory weight slice, so the update is a pure float4
// stream).  Pad rows / columns carry zero gradients and stay zero.
__device__ __forceinline__ void wgrad_tile(const float* __restrict__ A, int lda, int nq,
                                           const float* __restrict__ dY, int ldy, int B, int BS,
                                           float* __restrict__ wtile, float* __restrict__ mtile,
                                           float* __restrict__ vtile, const AdamC& ac,
                                           float* __restrict__ scratch) {
  constexpr int TC = kTpT;
  const int tid = threadIdx.x;
  const int n4 = TC * nq;                       // float4s of the tile (<= 2 per thread)
  float4 m4[2], v4[2];                          // moments: requested before the main loop
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int e = tid + i * kTpThreads;
    if (e < n4) {
      m4[i] = ldcg4(mtile + 4 * (size_t)e);
      v4[i] = ldcg4(vtile + 4 * (size_t)e);
    }
  }
  const int jq = tid % nq, bs = tid / nq;
  if (bs < BS) {
    const int rb = (B + BS - 1) / BS;
    const int b0 = bs * rb, b1 = min(B, b0 + rb);
    float acc[TC][4];
#pragma unroll
    for (int c = 0; c < TC; ++c)
#pragma unroll
      for (int u = 0; u < 4; ++u) acc[c][u] = 0.f;
    const float* ap = A + 4 * jq;
#pragma unroll 1
    for (int b = b0; b < b1; ++b) {
      const float4 a = *reinterpret_cast<const float4*>(ap + (size_t)b * lda);
#pragma unroll
      for (int t = 0; t < TC / 4; ++t) {
        const float4 d = *reinterpret_cast<const float4*>(dY + (size_t)b * ldy + 4 * t);
        fma4(acc[4 * t + 0], d.x, a);
        fma4(acc[4 * t + 1], d.y, a);
        fma4(acc[4 * t + 2], d.z, a);
        fma4(acc[4 * t + 3], d.w, a);
      }
    }
    float* s = scratch + 4 * ((size_t)bs * n4 + jq);
#pragma unroll
    for (int c = 0; c < TC; ++c)
      *reinterpret_cast<float4*>(s + 4 * (size_t)c * nq) =
          make_float4(acc[c][0], acc[c][1], acc[c][2], acc[c][3]);
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int e = tid + i * kTpThreads;
    if (e < n4) {
      float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
      for (int sidx = 0; sidx < BS; ++sidx) {
        const float4 t = *reinterpret_cast<const float4*>(scratch + 4 * ((size_t)sidx * n4 + e));
        g.x += t.x; g.y += t.y; g.z += t.z; g.w += t.w;
      }
      float4 w = *reinterpret_cast<float4*>(wtile + 4 * (size_t)e);
      adam1(w.x, g.x, m4[i].x, v4[i].x, ac);
      adam1(w.y, g.y, m4[i].y, v4[i].y, ac);
      adam1(w.z, g.z, m4[i].z, v4[i].z, ac);
      adam1(w.w, g.w, m4[i].w, v4[i].w, ac);
      *reinterpret_cast<float4*>(wtile + 4 * (size_t)e) = w;
      __stcg(reinterpret_cast<float4*>(mtile + 4 * (size_t)e), m4[i]);
      __stcg(reinterpret_cast<float4*>(vtile + 4 * (size_t)e), v4[i]);
    }
  }
  __syncthreads();
}

// ------------------------------------------------------------------------- dgrad partial
// part[b][j] = sum_{c < ldy} dY[b][c] * W[c][j] for every j < K (K % 4 == 0): the
// contribution of this CTA's columns to the gradient of the layer's input.
// threads <-> (j-quad, row group); rows of a group stride by the number of groups.
__device__ __forceinline__ void dgrad_partial(const float* __restrict__ dY, int ldy,
                                              const float* __restrict__ W, int ldw, int K,
                                              int B, float* __restrict__ out) {
  const int tid = threadIdx.x;
  const int nq = K >> 2;
  const int RG = kTpThreads / nq;                    // row groups
  const int jq = tid % nq, rg = tid / nq;
  if (rg >= RG) return;
  constexpr int TR = 4;                              // rows per pass
#pragma unroll 1
  for (int r0 = rg; r0 < B; r0 += RG * TR) {
    float acc[TR][4];
#pragma unroll
    for (int i = 0; i < TR; ++i)
#pragma unroll
      for (int u = 0; u < 4; ++u) acc[i][u] = 0.f;
#pragma unroll 1
    for (int c = 0; c < ldy; c += 4) {
      float4 w[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        w[u] = *reinterpret_cast<const float4*>(W + (size_t)(c + u) * ldw + 4 * jq);
#pragma unroll
      for (int i = 0; i < TR; ++i) {
        const int r = min(r0 + i * RG, B - 1);
        const float4 d = *reinterpret_cast<const float4*>(dY + (size_t)r * ldy + c);
        fma4(acc[i], d.x, w[0]);
        fma4(acc[i], d.y, w[1]);
        fma4(acc[i], d.z, w[2]);
        fma4(acc[i], d.w, w[3]);
      }
    }
#pragma unroll
    for (int i = 0; i < TR; ++i)
      if (r0 + i * RG < B)
        __stcg(reinterpret_cast<float4*>(out + (size_t)(r0 + i * RG) * K + 4 * jq),
               make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
  }
}

// ------------------------------------------------------------ mixture NLL, diagonal covariance
// One WARP per sample: lane = (ph, k) with k < GW the mixture component (GW = power of two
// >= K, run-time) and the P output dimensions dealt over the 32 / GW "p-lanes" (i = ph,
// ph + PS, ...).  Reads the staged head-output row z (smem), writes d loss / d z into dz
// (smem) WITHOUT the batch-global eps-gradient term (added by the owner of each column once
// the batch sum is known).  The forward parks L_d and the standardised residual in the dz
// row; the backward turns them into gradients in place.  Arithmetic of nll_small_kernel.
__device__ __forceinline__ float grp_max(float v, int gw) {
#pragma unroll 1
  for (int o = gw >> 1; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float grp_sum(float v, int gw) {
#pragma unroll 1
  for (int o = gw >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ void nll_rows_diag(const float* __restrict__ zr, float* __restrict__ dzr,
                                              int NHp, const float* __restrict__ nz,
                                              const float* __restrict__ ys, int nrow, int P, int K,
                                              int GW, float eps, float inv_b, float& loss_acc,
                                              float& s_acc, bool& bad) {
  const int tid = threadIdx.x, lane = tid & 31;
  const int PS = 32 / GW;
  const int k = lane & (GW - 1), ph = lane / GW;
  const int PK = P * K;
#pragma unroll 1
  for (int b = tid >> 5; b < nrow; b += kTpWarps) {
    const bool ok = k < K;
    const int kk = ok ? k : 0;
    const float* z = zr + (size_t)b * NHp;
    float* dz = dzr + (size_t)b * NHp;
    const float* nzr = nz + (size_t)b * PK;
    const float* yr = ys + (size_t)b * P;
    // mixture weights: softmax -> clamp -> renormalise (mdnn.py:109-111)
    const float zpi = ok ? z[kk] : -INFINITY;
    float mx = grp_max(zpi, GW);
    float soft = (zpi == -INFINITY) ? 0.f : expf(zpi - mx);
    const float sme = grp_sum(soft, GW);
    soft = soft / sme;
    float w = ok ? fminf(fmaxf(soft, kMinWeight), 1.0f) : 0.f;
    const float csum = grp_sum(w, GW);
    w = w / csum;
    // log density of component k: partial sums over this lane's dimensions
    float quad = 0.f, logdet = 0.f;
#pragma unroll 1
    for (int i = ph; i < P; i += PS) {
      const int o = i * K + kk;
      const float ldv = expf(z[K + PK + o]) + nzr[o] * eps;
      const float mu = z[K + o];
      bad |= ok && !(finite_f(ldv) && finite_f(mu));
      const float zi = (yr[i] - mu) / ldv;
      quad = fmaf(zi, zi, quad);
      logdet += logf(ldv);
      if (ok) { dz[K + o] = zi; dz[K + PK + o] = ldv; }
    }
#pragma unroll 1
    for (int o = GW; o < 32; o <<= 1) {
      quad += __shfl_xor_sync(0xffffffffu, quad, o);
      logdet += __shfl_xor_sync(0xffffffffu, logdet, o);
    }
    const float gj = -0.5f * ((float)P * kLog2Pi + quad) - logdet;
    bad |= ok && !(finite_f(gj) && finite_f(w));
    const float wc = fminf(fmaxf(w, kMinWeight), 1.0f);
    const float rk = ok ? fminf(fmaxf(gj, -kLLLimit), kLLLimit) + logf(wc) : -INFINITY;
    mx = grp_max(rk, GW);
    const float se = grp_sum(rk == -INFINITY ? 0.f : expf(rk - mx), GW);
    const float lse = mx + logf(se);
    if (lane == 0) loss_acc -= lse;
    // backward
    const float coef = ok ? -expf(rk - lse) * inv_b : 0.f;
    const bool in_w = (w >= kMinWeight) && (w <= 1.0f);
    const float dw = (ok && in_w) ? coef / wc : 0.f;
    const float t1 = grp_sum(dw * w, GW);
    const float dc = (dw - t1) / csum;
    const float dp = ((soft >= kMinWeight) && (soft <= 1.0f)) ? dc : 0.f;
    const float t2 = grp_sum(dp * soft, GW);
    if (ok && ph == 0) dz[k] = soft * (dp - t2);
    const float cg = ((gj >= -kLLLimit) && (gj <= kLLLimit)) ? coef : 0.f;
    __syncwarp();
#pragma unroll 1
    for (int i = ph; i < P; i += PS) {
      const int o = i * K + kk;
      const float ldv = dz[K + PK + o];
      const float zi = dz[K + o];
      const float nv = nzr[o];
      const float inv = 1.0f / ldv;
      const float vi = zi * inv;
      const float dld = cg * fmaf(vi, zi, -inv);
      if (ok) {
        dz[K + o] = cg * vi;
        s_acc = fmaf(dld, nv, s_acc);
        dz[K + PK + o] = fmaf(-nv, eps, ldv) * dld;        // exp(z_d) * d L_d
      }
    }
  }
}

// ======================================================================== the kernel
template <bool FULL, bool PROF>
__global__ void __launch_bounds__(kTpThreads, 1) train_persistent_kernel(const TpArgs a) {
  extern __shared__ __align__(16) float sm[];
  __shared__ float bsum[33];
  __shared__ float bc[4];
  __shared__ __align__(8) uint64_t xbar;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  unsigned rank_u;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank_u));
  const int rank = (int)rank_u;
  const int B = a.B, P = a.P, Kc = a.Kc, PK = P * Kc, NL = a.n_layers, NHp = a.NHp;
  float* RH = sm + a.rh;
  float* RX = sm + a.rx;
  const TpLayer& head = a.layer[NL - 1];
  float* mbase = a.gx + a.g_mv + (size_t)rank * 2 * a.wslice;   // this CTA's Adam moments
  float* vbase = mbase + a.wslice;

  // ---- one-time: weight slices (zero padded) into shared memory
  for (int l = 0; l < NL; ++l) {
    const TpLayer& L = a.layer[l];
    const int n0 = rank * L.cpc;
    for (int c = warp; c < L.ldy; c += kTpWarps) {
      const bool live = c < L.cpc && n0 + c < L.N;
      const float* src = a.params + L.w_off + (long long)(n0 + c) * L.K;
      for (int k = lane; k < L.ldw; k += 32)
        sm[L.sw + c * L.ldw + k] = (live && k < L.K) ? src[k] : 0.f;
      if (lane == 0) sm[L.sb + c] = live ? a.params[L.b_off + n0 + c] : 0.f;
    }
    for (int e = tid; e < B * L.ldy; e += kTpThreads) sm[L.own + e] = 0.f;
  }
  if (tid == 0) {
    ac::mbar_init(&xbar, 1);
    ac::fence_barrier_init();
  }
  __syncthreads();
  uint32_t xphase = 0;
  // optional per-phase cycle accounting (thread 0 of CTA 0; slots: profiles/tp_profile.py)
  long long tprev = 0;
  if (PROF) tprev = clock64();
  auto mark = [&](int slot) {
    if (PROF) {
      if (rank == 0 && tid == 0) {
        const long long t = clock64();
        a.prof[slot] += t - tprev;
        tprev = t;
      }
    }
  };
  // bulk-copy the gathered minibatch rows of update `u` into R_X (one copy per row)
  auto issue_x = [&](int u) {
    ac::fence_proxy_async();           // earlier generic-proxy accesses of R_X before the bulk writes
    if (tid == 0) ac::mbar_expect_tx(&xbar, (uint32_t)B * (uint32_t)a.ldx * 4u);
    __syncthreads();
    if (tid < B) {
      const long long row = a.idx[(long long)u * B + tid];
      ac::bulk_g2s(RX + (size_t)tid * a.ldxs, a.x + row * a.ldx, (uint32_t)a.ldx * 4u, &xbar);
    }
  };
  auto wait_x = [&]() {
    ac::mbar_wait(&xbar, xphase);
    xphase ^= 1u;
  };
  // gathered activation [B][N] (global, L2) -> R_H [B][ld]; four loads in flight per thread
  auto load_act = [&](const float* g, int N, int ld) {
    const int n4 = N >> 2;
    const int q = tid % n4, r0 = tid / n4, dr = kTpThreads / n4;
    if (r0 < dr) {
      const float* gp = g + 4 * q;
      float* dp = RH + 4 * q;
#pragma unroll 1
      for (int r = r0; r < B; r += 4 * dr) {
        float4 v[4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (r + i * dr < B) v[i] = ldcg4(gp + (size_t)(r + i * dr) * N);
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (r + i * dr < B) *reinterpret_cast<float4*>(dp + (size_t)(r + i * dr) * ld) = v[i];
      }
    }
  };
  cluster_arrive();
  cluster_wait();
  if (a.step0 < a.step1) issue_x(a.step0);

  AdamC adc;
  adc.one_minus_b1 = a.one_minus_b1; adc.b2 = a.b2; adc.one_minus_b2 = a.one_minus_b2; adc.eps = a.eps;

#pragma unroll 1
  for (int u = a.step0; u < a.step1; ++u) {
    adc.step_size = a.adam_coef[2 * u];
    adc.inv_bc2 = a.adam_coef[2 * u + 1];
    // ================================================================= forward
    mark(31);
    wait_x();
    mark(0);
#pragma unroll 1
    for (int l = 0; l < NL; ++l) {
      const TpLayer& L = a.layer[l];
      const bool is_head = (l == NL - 1);
      const float* A = (l == 0) ? RX : RH;
      float* scratch = (l == 0) ? RH : RX;
      const int n0 = rank * L.cpc;
      const int ncol = max(0, min(L.cpc, L.N - n0));
      const int ldg = is_head ? NHp : L.N;
      float* gout = a.gx + L.g_act + n0;
      const int e_lo = is_head ? Kc + PK : 0, e_hi = is_head ? Kc + 2 * PK : 0;
      float esum = 0.f;
#pragma unroll 1
      for (int c0 = 0; c0 < ncol; c0 += kTpT)
        esum += fwd_tile(A, L.lda, B, sm + L.sw + c0 * L.ldw, L.ldw, sm + L.sb + c0,
                         min(ncol - c0, kTpT), !is_head, sm + L.own + c0, L.ldy, gout + c0, ldg,
                         n0 + c0, e_lo, e_hi, scratch);
      if (is_head) {
        esum = block_sum(esum, bsum);
        if (tid == 0) __stcg(a.gx + a.g_sum + rank, esum);
      }
      mark(1 + 2 * l);
      cluster_arrive();
      cluster_wait();
      if (!is_head) load_act(a.gx + L.g_act, L.N, a.layer[l + 1].lda);
      __syncthreads();
      mark(2 + 2 * l);
    }

    // ============================================================ mixture NLL (row parallel)
    float* nb = (NL == 1) ? RH : RX;                 // phase-local buffers
    {
      const int row0 = rank * a.rpc;
      const int nrow = max(0, min(a.rpc, B - row0));
      float* zr = nb;                                // [rpc][NHp]
      float* dzr = zr + a.rpc * NHp;                 // [rpc][NHp]
      float* nz = dzr + a.rpc * NHp;                 // [rpc][PK]
      float* ys = nz + ((a.rpc * PK + 3) & ~3);      // [rpc][P]
      const float* zg = a.gx + head.g_act;
      for (int e = tid; e < nrow * (NHp >> 2); e += kTpThreads)
        reinterpret_cast<float4*>(zr)[e] = ldcg4(zg + (size_t)row0 * NHp + 4 * e);
      for (int e = tid; e < nrow * PK; e += kTpThreads)
        nz[e] = __ldg(a.noise + ((long long)u * B + row0) * PK + e);
      if (tid < nrow * P) {
        const int r = tid / P, i = tid - r * P;
        ys[tid] = __ldg(a.y + a.idx[(long long)u * B + row0 + r] * P + i);
      }
      if (tid < 32) {
        float v = tid < kTpNC ? __ldcg(a.gx + a.g_sum + tid) : 0.f;
        v = warp_sum(v);
        if (tid == 0) bc[0] = v;
      }
      __syncthreads();
      mark(7);
      const float eps = kEpsNoise * (bc[0] / (float)((long long)B * PK));
      float loss_acc = 0.f, s_acc = 0.f;
      bool bad = false;
      if (FULL) {
        float* fs = ys + ((a.rpc * P + 3) & ~3);     // 2 x [P][threads]
        NllArgs t;
        t.B = nrow; t.P = P; t.K = Kc; t.L = a.L;
        t.z_pi = zr; t.ld_pi = NHp;
        t.mu = zr + Kc; t.ld_mu = NHp;
        t.zd = zr + Kc + PK; t.ld_zd = NHp;
        t.low = zr + Kc + 2 * PK; t.ld_low = NHp;
        t.noise = nz;
        t.y = ys; t.y_rows = nullptr;
        t.grad_scale = nullptr; t.loss = nullptr; t.ws = nullptr; t.flag = a.flag; t.nparts_e = 0;
        t.d_pi = dzr; t.ldo_pi = NHp;
        t.d_mu = dzr + Kc; t.ldo_mu = NHp;
        t.d_zd = dzr + Kc + PK; t.ldo_zd = NHp;
        t.d_low = dzr + Kc + 2 * PK; t.ldo_low = NHp;
        float* zs = fs + tid;
        float* vs = fs + (size_t)P * kTpThreads + tid;
        const float ib = 1.0f / (float)B;
        if (a.gw <= 4)
          nll_samples<4, 1, true, true, true, true, false>(t, eps, ib, 0, kTpThreads / 4, kTpThreads, zs, vs, loss_acc, s_acc, bad);
        else if (a.gw == 8)
          nll_samples<8, 1, true, true, true, true, false>(t, eps, ib, 0, kTpThreads / 8, kTpThreads, zs, vs, loss_acc, s_acc, bad);
        else if (a.gw == 16)
          nll_samples<16, 1, true, true, true, true, false>(t, eps, ib, 0, kTpThreads / 16, kTpThreads, zs, vs, loss_acc, s_acc, bad);
        else
          nll_samples<32, 1, true, true, true, true, false>(t, eps, ib, 0, kTpThreads / 32, kTpThreads, zs, vs, loss_acc, s_acc, bad);
      } else {
        nll_rows_diag(zr, dzr, NHp, nz, ys, nrow, P, Kc, a.gw, eps, 1.0f / (float)B, loss_acc,
                      s_acc, bad);
      }
      if (bad) atomicOr(a.flag, 1);
      const float lsum = block_sum(loss_acc, bsum);
      const float ssum = block_sum(s_acc, bsum);
      if (tid == 0) {
        __stcg(a.gx + a.g_sum + kTpNC + rank, lsum);
        __stcg(a.gx + a.g_sum + 2 * kTpNC + rank, ssum);
      }
      float* dzg = a.gx + a.g_dz;
      for (int e = tid; e < nrow * (NHp >> 2); e += kTpThreads)
        __stcg(reinterpret_cast<float4*>(dzg + (size_t)row0 * NHp) + e,
               reinterpret_cast<const float4*>(dzr)[e]);
    }
    mark(8);
    cluster_arrive();
    cluster_wait();
    {
      // own columns of dz (+ the batch-global eps-gradient term on the log-diag columns):
      // overwrites the own-column copy of z
      if (tid < 32) {
        float l = tid < kTpNC ? __ldcg(a.gx + a.g_sum + kTpNC + tid) : 0.f;
        float s = tid < kTpNC ? __ldcg(a.gx + a.g_sum + 2 * kTpNC + tid) : 0.f;
        l = warp_sum(l);
        s = warp_sum(s);
        if (tid == 0) { bc[1] = l; bc[2] = s; }
      }
      __syncthreads();
      const int slot = a.loss_slot[u];
      if (rank == 0 && tid == 0 && slot >= 0) a.loss[slot] = bc[1] / (float)B;
      const float cfix = kEpsNoise * bc[2] / (float)((long long)B * PK);
      const int n0 = rank * head.cpc;
      const int ncol = max(0, min(head.cpc, head.N - n0));
      const float* dzg = a.gx + a.g_dz + n0;
      const int r = tid & 127;                       // thread <-> (row, column mod 4)
      if (r < B) {
#pragma unroll 1
        for (int c = tid >> 7; c < head.ldy; c += 4) {
          float d = 0.f;
          if (c < ncol) {
            const int col = n0 + c;
            d = __ldcg(dzg + (size_t)r * NHp + c);
            if (col >= Kc + PK && col < Kc + 2 * PK) d += expf(sm[head.own + r * head.ldy + c]) * cfix;
          }
          sm[head.own + r * head.ldy + c] = d;
        }
      }
      __syncthreads();
      mark(9);
    }

    // ================================================================ backward
#pragma unroll 1
    for (int l = NL - 1; l >= 0; --l) {
      const TpLayer& L = a.layer[l];
      const int n0 = rank * L.cpc;
      const int ncol = max(0, min(L.cpc, L.N - n0));
      const int pb = 10 + 4 * (NL - 1 - l);
      if (l > 0) {
        dgrad_partial(sm + L.own, L.ldy, sm + L.sw, L.ldw, L.K, B,
                      a.gx + L.g_part + (size_t)rank * B * L.K);
        mark(pb);
        cluster_arrive();
      }
      // ---- weight gradient of the own columns + Adam (local)
      if (l == 0 && NL > 1) wait_x();
      mark(pb + 1);
      {
        const float* A = (l == 0) ? RX : RH;
        float* scratch = (l == 0) ? RH : RX;
        const int cap = (l == 0) ? a.rh_floats : a.rx_floats;
        const int nq = L.ldw >> 2;
        const int BS = max(1, min(min(kTpThreads / nq, cap / (kTpT * L.ldw)), B));
#pragma unroll 1
        for (int c0 = 0; c0 < ncol; c0 += kTpT)
          wgrad_tile(A, L.lda, nq, sm + L.own + c0, L.ldy, B, BS, sm + L.sw + c0 * L.ldw,
                     mbase + L.sw + c0 * L.ldw, vbase + L.sw + c0 * L.ldw, adc, scratch);
        // bias gradient = column sums of dY: one warp per column
#pragma unroll 1
        for (int c = warp; c < ncol; c += kTpWarps) {
          float g = 0.f;
          for (int b = lane; b < B; b += 32) g += sm[L.own + b * L.ldy + c];
          g = warp_sum(g);
          if (lane == 0) {
            float mm = __ldcg(mbase + L.sb + c), vv = __ldcg(vbase + L.sb + c);
            adam1(sm[L.sb + c], g, mm, vv, adc);
            __stcg(mbase + L.sb + c, mm);
            __stcg(vbase + L.sb + c, vv);
          }
        }
        __syncthreads();
      }
      mark(pb + 2);
      if (l == 0) {
        // x of the next update can start to arrive (R_X is free: wgrad 0 has read it)
        if (u + 1 < a.step1) issue_x(u + 1);
      } else {
        // the input of the next wgrad: hidden activation l-2 back into R_H / x back into R_X
        if (l >= 2) load_act(a.gx + a.layer[l - 2].g_act, a.layer[l - 2].N, a.layer[l - 1].lda);
        else issue_x(u);
        cluster_wait();
        // reduce-scatter: own columns of d h_{l-1} times tanh' (over the own-column copy of h)
        const TpLayer& Lp = a.layer[l - 1];
        const int pn0 = rank * Lp.cpc;
        const int pncol = max(0, min(Lp.cpc, Lp.N - pn0));
        const float* part = a.gx + L.g_part + pn0;
        if ((Lp.cpc & 3) == 0) {
          // whole float4s are valid (N % 4 == 0, cpc % 4 == 0): 16 vector loads in flight
          const int q4 = Lp.ldy >> 2;
#pragma unroll 1
          for (int e = tid; e < B * q4; e += kTpThreads) {
            const int r = e / q4, c = 4 * (e - r * q4);
            float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c < pncol) {
              float4 v[kTpNC];
#pragma unroll
              for (int q = 0; q < kTpNC; ++q) v[q] = ldcg4(part + ((size_t)q * B + r) * L.K + c);
#pragma unroll
              for (int q = 0; q < kTpNC; ++q) { s.x += v[q].x; s.y += v[q].y; s.z += v[q].z; s.w += v[q].w; }
              const float4 h = *reinterpret_cast<const float4*>(sm + Lp.own + r * Lp.ldy + c);
              s.x *= (1.f - h.x * h.x); s.y *= (1.f - h.y * h.y);
              s.z *= (1.f - h.z * h.z); s.w *= (1.f - h.w * h.w);
            }
            *reinterpret_cast<float4*>(sm + Lp.own + r * Lp.ldy + c) = s;
          }
        } else {
          const int r = tid & 127;                   // thread <-> (row, column mod 4)
          if (r < B) {
#pragma unroll 1
            for (int c = tid >> 7; c < Lp.ldy; c += 4) {
              float s = 0.f;
              if (c < pncol) {
                float v[kTpNC];
#pragma unroll
                for (int q = 0; q < kTpNC; ++q) v[q] = __ldcg(part + ((size_t)q * B + r) * L.K + c);
#pragma unroll
                for (int q = 0; q < kTpNC; ++q) s += v[q];
                const float h = sm[Lp.own + r * Lp.ldy + c];
                s *= (1.f - h * h);
              }
              sm[Lp.own + r * Lp.ldy + c] = s;
            }
          }
        }
        __syncthreads();
        mark(pb + 3);
      }
    }
  }

  // ---- write the parameter slices back
  __syncthreads();
  for (int l = 0; l < NL; ++l) {
    const TpLayer& L = a.layer[l];
    const int n0 = rank * L.cpc;
    for (int c = warp; c < L.cpc; c += kTpWarps) {
      if (n0 + c >= L.N) continue;
      float* dst = a.params + L.w_off + (long long)(n0 + c) * L.K;
      for (int k = lane; k < L.K; k += 32) dst[k] = sm[L.sw + c * L.ldw + k];
      if (lane == 0) a.params[L.b_off + n0 + c] = sm[L.sb + c];
    }
  }
  cluster_arrive();
  cluster_wait();
}

// --------------------------------------------------------------------------- host side
static int pad_pitch(int k) {      // multiple of 4 with pitch mod 32 in {4, 12, 20, 28}
  int p = (k + 3) & ~3;
  while ((p & 7) != 4) p += 4;
  return p;
}
static int tiles_pad(int cpc) {    // column tiles are 8 wide
  return (cpc + kTpT - 1) / kTpT * kTpT;
}

static int tp_plan(const bsig_tp_desc* d, TpArgs& a, long long* scratch_floats, long long* smem_bytes) {
  BSIG_REQUIRE(d != nullptr, "train_persistent: null descriptor");
  const int NL = d->n_layers;
  BSIG_REQUIRE(NL >= 1 && NL <= kTpMaxLayers, "train_persistent: 1..3 dense layers (got %d)", NL);
  const int B = d->batch, P = d->p, K = d->k;
  BSIG_REQUIRE(B >= 1 && B <= 128, "train_persistent: minibatch 1..128 (got %d)", B);
  BSIG_REQUIRE(K >= 1 && K <= 32 && P >= 1, "train_persistent: 1..32 mixture components");
  const int Lsz = d->full_cov ? P * (P - 1) / 2 : 0;
  const int NH = K * (1 + 2 * P + Lsz);
  BSIG_REQUIRE(d->out_dim[NL - 1] == NH, "train_persistent: head width %d != K(1+2P+L) = %d",
               d->out_dim[NL - 1], NH);
  a = TpArgs();
  a.B = B; a.F = d->in_dim[0]; a.P = P; a.Kc = K; a.L = Lsz; a.NH = NH; a.NHp = (NH + 3) & ~3;
  a.n_layers = NL;
  a.gw = 1;
  while (a.gw < K) a.gw <<= 1;
  a.rpc = (B + kTpNC - 1) / kTpNC;
  BSIG_REQUIRE(d->ldx % 4 == 0 && d->ldx >= a.F, "train_persistent: x pitch must be a multiple of 4 floats");
  int off = 0;
  long long goff = 0;
  int rh_need = 0, rx_need = 0;
  // shared memory: [weight + bias slices of all layers] [own / dy copies] [R_H] [R_X]
  for (int l = 0; l < NL; ++l) {
    TpLayer& L = a.layer[l];
    L.K = d->in_dim[l]; L.N = d->out_dim[l];
    BSIG_REQUIRE(L.K >= 1 && L.N >= 1, "train_persistent: bad layer shape");
    if (l > 0) BSIG_REQUIRE(L.K == a.layer[l - 1].N, "train_persistent: layer widths do not chain");
    if (l < NL - 1) BSIG_REQUIRE(L.N % 4 == 0, "train_persistent: hidden width %d must be a multiple of 4", L.N);
    L.cpc = (L.N + kTpNC - 1) / kTpNC;
    L.ldy = tiles_pad(L.cpc);
    L.ldw = (L.K + 3) & ~3;
    L.lda = (l == 0) ? pad_pitch((int)d->ldx) : pad_pitch(L.K);
    L.w_off = d->w_off[l]; L.b_off = d->b_off[l];
    BSIG_REQUIRE(kTpT * (L.ldw >> 2) <= 2 * kTpThreads && (L.ldw >> 2) <= kTpThreads,
                 "train_persistent: layer %d (%d -> %d) too wide for one cluster", l, L.K, L.N);
    if (l > 0) BSIG_REQUIRE((L.K >> 2) <= kTpThreads, "train_persistent: hidden width too large");
    L.sw = off; off += L.ldy * L.ldw;
    L.sb = off; off += L.ldy;
  }
  a.wslice = off;
  for (int l = 0; l < NL; ++l) {
    TpLayer& L = a.layer[l];
    L.own = off; off += B * L.ldy;
    L.g_act = goff; goff += (long long)B * ((l == NL - 1) ? a.NHp : L.N);
    L.g_part = goff; if (l > 0) goff += (long long)kTpNC * B * L.K;
    const int fwd_scr = kTpWarps * kTpT * B;
    const int wg_scr = kTpT * L.ldw;                   // at least one batch split
    if (l == 0) rh_need = std::max(rh_need, std::max(fwd_scr, wg_scr));
    else rx_need = std::max(rx_need, std::max(fwd_scr, wg_scr));
    if (l > 0) rh_need = std::max(rh_need, B * L.lda);
  }
  a.g_dz = goff; goff += (long long)B * a.NHp;
  a.g_sum = goff; goff += 4 * kTpNC;
  a.g_mv = goff; goff += (long long)kTpNC * 2 * a.wslice;
  a.ldxs = a.layer[0].lda;
  rx_need = std::max(rx_need, B * a.ldxs);
  const int nll_need = 2 * a.rpc * a.NHp + ((a.rpc * P * K + 3) & ~3) + ((a.rpc * P + 3) & ~3) +
                       (d->full_cov ? 2 * P * kTpThreads : 0) + 16;
  if (NL == 1) rh_need = std::max(rh_need, nll_need); else rx_need = std::max(rx_need, nll_need);
  BSIG_REQUIRE(a.rpc * P <= kTpThreads, "train_persistent: too many output dimensions");
  a.rh = off; a.rh_floats = (rh_need + 3) & ~3; off += a.rh_floats;
  a.rx = off; a.rx_floats = (rx_need + 3) & ~3; off += a.rx_floats;
  const long long smem = (long long)off * 4;
  BSIG_REQUIRE(smem <= 227 * 1024 - 512, "train_persistent: needs %lld bytes of shared memory per CTA (> 227 KB)", smem);
  if (scratch_floats) *scratch_floats = goff;
  if (smem_bytes) *smem_bytes = smem;
  a.x = d->x; a.ldx = d->ldx; a.y = d->y; a.idx = (const long long*)d->idx; a.noise = d->noise;
  a.params = d->params; a.gx = d->scratch;
  a.loss = d->loss_buf; a.loss_slot = d->loss_slot; a.adam_coef = d->adam_coef; a.flag = d->flag;
  a.prof = (long long*)d->prof;
  a.one_minus_b1 = 1.0f - d->beta1; a.b2 = d->beta2; a.one_minus_b2 = 1.0f - d->beta2; a.eps = d->eps;
  return 0;
}

template <bool FULL, bool PROF>
static int tp_launch(const TpArgs& a, size_t smem, cudaStream_t st) {
  auto kern = train_persistent_kernel<FULL, PROF>;
  static size_t configured = 0;      // largest dynamic shared-memory size enabled so far
  if (configured == 0)
    BSIG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  if (smem > configured) {
    BSIG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(kTpNC);
  cfg.blockDim = dim3(kTpThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kTpNC;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  BSIG_CUDA(cudaLaunchKernelEx(&cfg, kern, a));
  BSIG_LAUNCH_CHECK();
  return 0;
}

}  // namespace bsig

extern "C" int bsig_train_persistent_query(const bsig_tp_desc* desc, int64_t* scratch_floats,
                                           int64_t* smem_bytes) {
  bsig::TpArgs a;
  long long sf = 0, sb = 0;
  const int rc = bsig::tp_plan(desc, a, &sf, &sb);
  if (rc != 0) return rc;
  if (scratch_floats) *scratch_floats = sf;
  if (smem_bytes) *smem_bytes = sb;
  return 0;
}

extern "C" int bsig_train_persistent(const bsig_tp_desc* desc, int64_t step0, int64_t step1,
                                     void* stream) {
  using namespace bsig;
  TpArgs a;
  long long sf = 0, sb = 0;
  const int rc = tp_plan(desc, a, &sf, &sb);
  if (rc != 0) return rc;
  BSIG_REQUIRE(desc->scratch != nullptr && desc->scratch_floats >= sf,
               "train_persistent: scratch too small (%lld < %lld floats)",
               (long long)desc->scratch_floats, sf);
  BSIG_REQUIRE(step0 >= 0 && step1 >= step0, "train_persistent: bad update range");
  if (step1 == step0) return 0;
  a.step0 = (int)step0; a.step1 = (int)step1;
  cudaStream_t st = (cudaStream_t)stream;
  const bool full = desc->full_cov != 0 && a.L > 0;
  const bool prof = a.prof != nullptr;
  if (full) return prof ? tp_launch<true, true>(a, (size_t)sb, st) : tp_launch<true, false>(a, (size_t)sb, st);
  return prof ? tp_launch<false, true>(a, (size_t)sb, st) : tp_launch<false, false>(a, (size_t)sb, st);
}
