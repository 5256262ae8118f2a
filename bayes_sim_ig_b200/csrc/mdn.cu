// Mixture-density head epilogue and mixture-of-Gaussians negative log-likelihood
// (reference bayes_sim_ig/models/mdnn.py:109-119 and 127-178), forward and
// backward, as explicit kernels.
//
// Layout of the concatenated head output z [B, NH]:
//   [0,K) pi logits | [K,K+PK) mu, col p*K+k | [K+PK,K+2PK) log-diag |
//   [K+2PK, K+2PK+LK) strict-lower entries, col l*K+k (np.tril_indices order)
//
// Work decomposition of the NLL kernels: a sub-warp group of GW lanes (GW =
// pow2 >= min(K,32)) owns one sample; lane g handles components g, g+GW, ...
// Per component the Cholesky-parameterised log-density needs a forward
// substitution (z = L^-1 (y-mu)), the backward pass a back substitution
// (v = L^-T z); the logsumexp over components is a shuffle reduction inside
// the group.  All of it is HBM/L2-bound streaming over z: each entry of z is
// read twice (forward, backward) and each entry of dz written once.
//
// Every cross-sample reduction (mean of exp(z_d) for the eps-noise term, the
// loss, the eps-term gradient) is two-stage and fixed-order => deterministic.
#include <cooperative_groups.h>

#include <algorithm>
#include <cstdlib>

#include "async_copy.cuh"
#include "common.cuh"

namespace cg = cooperative_groups;

#include "mdn_core.cuh"

namespace bsig {

// ---------------------------------------------------------------- exp-sum (eps)
__global__ void __launch_bounds__(256)
exp_sum_kernel(const float* __restrict__ zd, int64_t ld_zd, int B, int PK, float* ws) {
  __shared__ float scratch[33];
  float acc = 0.f;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t start = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (((PK | ld_zd) & 1) == 0 && (reinterpret_cast<uintptr_t>(zd) & 7) == 0) {
    // even widths: 64-bit loads (rows of the fused head output are only 8-byte aligned)
    const int W = PK >> 1;
    const int64_t total = (int64_t)B * W;
    RowCol rc(start, stride, W);
    for (int64_t i = start; i < total; i += stride, rc.next()) {
      const float2 v = __ldg(reinterpret_cast<const float2*>(zd + rc.row * ld_zd) + rc.col);
      acc += expf(v.x) + expf(v.y);
    }
  } else {
    const int64_t total = (int64_t)B * PK;
    RowCol rc(start, stride, PK);
    for (int64_t i = start; i < total; i += stride, rc.next())
      acc += expf(__ldg(zd + rc.row * ld_zd + rc.col));
  }
  acc = block_sum(acc, scratch);
  if (threadIdx.x == 0) ws[blockIdx.x] = acc;
}

__device__ __forceinline__ float sum_parts(const float* parts, int n, float* scratch) {
  float acc = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) acc += parts[i];
  return block_sum(acc, scratch);
}


template <int GW, int KPL, bool FUSED, bool FULL, bool BWD>
__global__ void __launch_bounds__(128) nll_kernel(NllArgs a) {
  extern __shared__ float dyn[];           // FULL: zs[P][128], vs[P][128]
  __shared__ float scratch[33];
  const int tid = threadIdx.x;
  constexpr int GPB = 128 / GW;
  const int B = a.B, P = a.P, K = a.K;
  float* zs = dyn + tid;                   // element p at zs[p*128]
  float* vs = dyn + (size_t)P * 128 + tid;

  float eps = 0.f;
  if (FUSED) {
    const float esum = sum_parts(a.ws, a.nparts_e, scratch);
    eps = kEpsNoise * (esum / (float)((int64_t)B * P * K));
  }
  float gscale = 1.f;
  if (BWD && a.grad_scale != nullptr) gscale = __ldg(a.grad_scale);
  const float coef_scale = gscale / (float)B;

  float loss_acc = 0.f, s_acc = 0.f;
  bool bad = false;
  nll_samples<GW, KPL, FUSED, FULL, BWD>(a, eps, coef_scale, blockIdx.x * GPB, gridDim.x * GPB,
                                         128, zs, vs, loss_acc, s_acc, bad);

  if (bad) atomicOr(a.flag, 1);

  // ---- deterministic cross-block reductions
  float* loss_parts = a.ws + kMaxParts;
  float* s_parts = a.ws + 2 * kMaxParts;
  unsigned int* counter = reinterpret_cast<unsigned int*>(a.ws + 3 * kMaxParts);
  const float lsum = block_sum(loss_acc, scratch);
  float ssum = 0.f;
  if (FUSED && BWD) ssum = block_sum(s_acc, scratch);
  __shared__ bool is_last;
  if (tid == 0) {
    loss_parts[blockIdx.x] = lsum;
    if (FUSED && BWD) s_parts[blockIdx.x] = ssum;
    __threadfence();
    const unsigned int done = atomicAdd(counter, 1u);
    is_last = (done == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    float acc = 0.f;
    for (int i = tid; i < (int)gridDim.x; i += blockDim.x) acc += __ldcg(loss_parts + i);
    acc = block_sum(acc, scratch);
    if (tid == 0) {
      a.loss[0] = acc / (float)B;
      *counter = 0u;   // ready for the next launch
    }
  }
}

// Large-batch streaming form of the fused head + NLL (forward, or forward + backward):
// persistent, warp-specialised CTAs walk tiles of R consecutive samples.
//   producer warp : per tile, one cp.async.bulk of the z tile ([R, NH], contiguous in
//                   HBM) and one of the eps-noise tile ([R, P*K]) into a double buffer
//                   (completion on an mbarrier), the (optionally gathered) y rows with
//                   plain loads, and -- once the consumers have finished a tile -- ONE
//                   bulk store of its dz back to HBM;
//   consumer warps: the per-sample routine, entirely out of shared memory (lane k of a
//                   sample reads column i*K + k: conflict-free), synchronised with the
//                   producer through mbarriers only (no CTA-wide barrier in the loop).
// With diagonal covariance dz is formed in place (a lane overwrites exactly the entries
// of z it has read); the full-covariance back substitution re-reads strict-lower entries
// of finished rows, so dz gets its own stage there.  HBM sees only full-line bulk
// traffic.  A ragged last tile (rows % 4 != 0: bulk copies need 16-byte multiples) is
// moved with plain loads / stores by the producer warp.
struct NllStream {
  const float* z;       // [B, NH]
  float* dz;            // [B, NH] (BWD)
  int NH, R;            // row width, rows per tile (multiple of 4)
};

template <int GW, int KPL, bool FULL, bool BWD>
__global__ void __launch_bounds__(288) nll_stream_kernel(NllArgs a, NllStream q) {
  extern __shared__ __align__(16) float dyn[];
  __shared__ float scratch[33];
  __shared__ __align__(8) uint64_t full[2], done[2];
  __shared__ bool is_last;
  const int tid = threadIdx.x, lane = tid & 31;
  const int TC = blockDim.x - 32;                 // consumer threads (warps 0 .. n-1)
  const bool producer = tid >= TC;
  const int B = a.B, P = a.P, K = a.K, PK = P * K, NH = q.NH, R = q.R;
  const int GPB = LaneGroup<GW>::per_cta(TC, K);
  constexpr bool STAGE = BWD && FULL;             // separate dz stage
  // per buffer: z[R*NH] | noise[R*PK] | y[R*P padded to 4] | (dz[R*NH]);  then zs, vs
  const int ypad = (R * P + 3) & ~3;
  const size_t bufsz = (size_t)R * NH + (size_t)R * PK + ypad + (STAGE ? (size_t)R * NH : 0);
  auto zbuf = [&](int b) { return dyn + b * bufsz; };
  auto nbuf = [&](int b) { return dyn + b * bufsz + (size_t)R * NH; };
  auto ybuf = [&](int b) { return dyn + b * bufsz + (size_t)R * NH + (size_t)R * PK; };
  auto dbuf = [&](int b) { return STAGE ? ybuf(b) + ypad : zbuf(b); };
  float* fs = dyn + 2 * bufsz;
  float* zs = fs + tid;
  float* vs = fs + (size_t)P * TC + tid;

  const float esum = sum_parts(a.ws, a.nparts_e, scratch);
  const float eps = kEpsNoise * (esum / (float)((int64_t)B * PK));
  const float coef_scale = 1.0f / (float)B;
  const int64_t ntiles = ((int64_t)B + R - 1) / R;
  // tiles of this CTA: blockIdx.x, blockIdx.x + gridDim.x, ...
  const int my_tiles = (int64_t)blockIdx.x < ntiles
                           ? (int)((ntiles - 1 - blockIdx.x) / gridDim.x) + 1 : 0;
  auto tile_of = [&](int it) { return (int64_t)blockIdx.x + (int64_t)it * gridDim.x; };
  auto tile_rows = [&](int64_t tile) { return (int)min((int64_t)R, (int64_t)B - tile * R); };

  if (tid == 0) {
    ac::mbar_init(&full[0], 2);                  // bulk copies (expect_tx) + y rows
    ac::mbar_init(&full[1], 2);
    ac::mbar_init(&done[0], TC >> 5);
    ac::mbar_init(&done[1], TC >> 5);
    ac::fence_barrier_init();
  }
  __syncthreads();

  float loss_acc = 0.f, s_acc = 0.f;
  bool bad = false;
  if (producer) {
    constexpr int YR = 8;                          // y values a lane keeps in flight
    for (int it = 0; it < my_tiles + 2; ++it) {
      const int b = it & 1;
      // y rows of tile `it`: request them now, park them in shared memory further down
      // (their latency hides behind the wait for the consumers)
      float yreg[YR];
      const bool load_tile = it < my_tiles;
      const int64_t ltile = load_tile ? tile_of(it) : 0;
      const int lrows = load_tile ? tile_rows(ltile) : 0;
      const bool y_regs = R * P <= YR * 32;
      if (load_tile && y_regs) {
#pragma unroll
        for (int u = 0; u < YR; ++u) {
          const int e = lane + u * 32;
          yreg[u] = 0.f;
          if (e < lrows * P) {
            const int r = e / P, c = e - r * P;
            const int64_t row = ltile * R + r;
            yreg[u] = __ldg(a.y + (a.y_rows ? __ldg(a.y_rows + row) : row) * P + c);
          }
        }
      }
      if (it >= 2) {
        // retire tile it-2: consumers are done with buffer b -> write its dz back
        const int64_t tile = tile_of(it - 2);
        const int rows = tile_rows(tile);
        ac::mbar_wait(&done[b], (uint32_t)((it - 2) >> 1) & 1u);
        if (BWD) {
          float* dzg = q.dz + tile * R * (int64_t)NH;
          if ((rows & 3) == 0) {
            if (lane == 0) {
              ac::bulk_s2g(dzg, dbuf(b), (uint32_t)rows * (uint32_t)NH * 4u);
              ac::bulk_commit();
              ac::bulk_wait_read<0>();           // buffer b is refilled below
            }
          } else {
            const float* src = dbuf(b);
            for (int e = lane; e < rows * NH; e += 32) dzg[e] = src[e];
          }
          __syncwarp();
        }
      }
      if (load_tile) {
        const float* zg = q.z + ltile * R * (int64_t)NH;
        const float* ng = a.noise + ltile * R * (int64_t)PK;
        if ((lrows & 3) == 0) {
          if (lane == 0) {
            ac::mbar_expect_tx(&full[b], (uint32_t)lrows * (uint32_t)(NH + PK) * 4u);
            ac::bulk_g2s(zbuf(b), zg, (uint32_t)lrows * NH * 4u, &full[b]);
            ac::bulk_g2s(nbuf(b), ng, (uint32_t)lrows * PK * 4u, &full[b]);
          }
        } else {
          float* zt = zbuf(b);
          float* nt = nbuf(b);
          for (int e = lane; e < lrows * NH; e += 32) zt[e] = __ldg(zg + e);
          for (int e = lane; e < lrows * PK; e += 32) nt[e] = __ldg(ng + e);
          __syncwarp();
          if (lane == 0) ac::mbar_arrive(&full[b]);
        }
        float* yt = ybuf(b);
        if (y_regs) {
#pragma unroll
          for (int u = 0; u < YR; ++u)
            if (lane + u * 32 < lrows * P) yt[lane + u * 32] = yreg[u];
        } else {
          for (int e = lane; e < lrows * P; e += 32) {
            const int r = e / P, c = e - r * P;
            const int64_t row = ltile * R + r;
            yt[e] = __ldg(a.y + (a.y_rows ? __ldg(a.y_rows + row) : row) * P + c);
          }
        }
        __syncwarp();
        if (lane == 0) ac::mbar_arrive(&full[b]);   // second arrival: y is in place
      }
    }
  } else {
    for (int it = 0; it < my_tiles; ++it) {
      const int b = it & 1;
      const int rows = tile_rows(tile_of(it));
      ac::mbar_wait(&full[b], (uint32_t)(it >> 1) & 1u);
      float* zt = zbuf(b);
      float* dzt = dbuf(b);
      NllArgs t = a;
      t.B = rows;
      t.z_pi = zt; t.ld_pi = NH;
      t.mu = zt + K; t.ld_mu = NH;
      t.zd = zt + K + PK; t.ld_zd = NH;
      t.low = FULL ? zt + K + 2 * PK : nullptr; t.ld_low = NH;
      t.noise = nbuf(b);
      t.y = ybuf(b); t.y_rows = nullptr;
      if (BWD) {
        t.d_pi = dzt; t.ldo_pi = NH;
        t.d_mu = dzt + K; t.ldo_mu = NH;
        t.d_zd = dzt + K + PK; t.ldo_zd = NH;
        t.d_low = FULL ? dzt + K + 2 * PK : nullptr; t.ldo_low = NH;
      }
      nll_samples<GW, KPL, true, FULL, BWD, true>(t, eps, coef_scale, 0, GPB, TC, zs, vs,
                                                  loss_acc, s_acc, bad);
      if (BWD) ac::fence_proxy_async();    // dz (generic writes) -> visible to the bulk store
      __syncwarp();
      if (lane == 0) ac::mbar_arrive(&done[b]);
    }
  }
  if (bad) atomicOr(a.flag, 1);

  // ---- deterministic cross-block reductions (same scheme as nll_kernel)
  float* loss_parts = a.ws + kMaxParts;
  float* s_parts = a.ws + 2 * kMaxParts;
  unsigned int* counter = reinterpret_cast<unsigned int*>(a.ws + 3 * kMaxParts);
  const float lsum = block_sum(loss_acc, scratch);
  const float ssum = BWD ? block_sum(s_acc, scratch) : 0.f;
  if (tid == 0) {
    loss_parts[blockIdx.x] = lsum;
    if (BWD) s_parts[blockIdx.x] = ssum;
    __threadfence();
    const unsigned int done_ctas = atomicAdd(counter, 1u);
    is_last = (done_ctas == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    float acc = 0.f;
    for (int i = tid; i < (int)gridDim.x; i += blockDim.x) acc += __ldcg(loss_parts + i);
    acc = block_sum(acc, scratch);
    if (tid == 0) {
      a.loss[0] = acc / (float)B;
      *counter = 0u;
    }
  }
}

// Sum one float per CTA over the cluster (call after cluster.sync()): the first
// warp gathers the NC partials through DSMEM in parallel and adds them in a fixed
// butterfly order, so every CTA obtains the bit-identical total.
__device__ __forceinline__ float cluster_sum(cg::cluster_group& cluster, float* slot, int NC,
                                             float* bcast) {
  if (threadIdx.x < 32) {
    float v = ((int)threadIdx.x < NC) ? *cluster.map_shared_rank(slot, threadIdx.x) : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) *bcast = v;
  }
  __syncthreads();
  return *bcast;
}

// (SFU transcendentals -- ex2 / lg2 / rcp.approx, <= 2 ulp -- as in the streaming kernel: the
// minibatch step is a chain of dependent instructions and the full-precision expf / logf /
// division sequences were a third of them)
// Small-batch form (B*GW <= 8*512 lanes, i.e. the reference's minibatch of 100
// and its 200-row test split): exp-sum, NLL forward/backward, eps-term fix-up and
// the loss in ONE launch.  The CTAs form a single thread-block cluster; the three
// batch-wide reductions go through distributed shared memory in rank order.
template <int GW, int KPL, bool FULL, bool BWD>
__global__ void __launch_bounds__(512) nll_cluster_kernel(NllArgs a) {
  extern __shared__ float dyn[];           // FULL: zs[P][TS], vs[P][TS]
  __shared__ float scratch[33];
  __shared__ float red[3];                 // this CTA's exp-sum, loss, eps-gradient partials
  cg::cluster_group cluster = cg::this_cluster();
  const int tid = threadIdx.x, TS = blockDim.x;
  const int NC = gridDim.x, rank = blockIdx.x;
  const int GPB = TS / GW;
  const int B = a.B, P = a.P, K = a.K, PK = P * K;
  float* zs = dyn + tid;
  float* vs = dyn + (size_t)P * TS + tid;
  pdl_wait_then_release();

  // phase 0: sum of exp(z_d) over the whole batch
  {
    const int64_t total = (int64_t)B * PK;
    const int64_t stride = (int64_t)NC * TS, start = (int64_t)rank * TS + tid;
    RowCol rc(start, stride, PK);
    float acc = 0.f;
    for (int64_t i = start; i < total; i += stride, rc.next())
      acc += expf(__ldg(a.zd + rc.row * a.ld_zd + rc.col));
    acc = block_sum(acc, scratch);
    if (tid == 0) red[0] = acc;
  }
  cluster.sync();
  float esum = 0.f;
  for (int r = 0; r < NC; ++r) esum += *cluster.map_shared_rank(&red[0], r);
  const float eps = kEpsNoise * (esum / (float)((int64_t)B * PK));
  const float coef_scale = 1.0f / (float)B;

  // phase 1: per-sample forward (+ backward)
  float loss_acc = 0.f, s_acc = 0.f;
  bool bad = false;
  nll_samples<GW, KPL, true, FULL, BWD>(a, eps, coef_scale, rank * GPB, NC * GPB, TS, zs, vs,
                                        loss_acc, s_acc, bad);
  if (bad) atomicOr(a.flag, 1);
  const float lsum = block_sum(loss_acc, scratch);
  const float ssum = BWD ? block_sum(s_acc, scratch) : 0.f;
  if (tid == 0) { red[1] = lsum; red[2] = ssum; }
  cluster.sync();
  if (rank == 0 && tid == 0) {
    float acc = 0.f;
    for (int r = 0; r < NC; ++r) acc += *cluster.map_shared_rank(&red[1], r);
    a.loss[0] = acc / (float)B;
  }
  if (BWD) {
    // phase 2: eps-term fix-up of the samples this CTA owns
    float S = 0.f;
    for (int r = 0; r < NC; ++r) S += *cluster.map_shared_rank(&red[2], r);
    const float c = kEpsNoise * S / (float)((int64_t)B * PK);
    for (int base = rank * GPB; base < B; base += NC * GPB) {
      for (int e = tid; e < GPB * PK; e += TS) {
        const int b = base + e / PK, col = e % PK;
        if (b < B)
          a.d_zd[(int64_t)b * a.ldo_zd + col] += expf(__ldg(a.zd + (int64_t)b * a.ld_zd + col)) * c;
      }
    }
  }
  cluster.sync();       // keep red[] alive until every CTA has read it
}

// Register-resident minibatch form (diagonal covariance, one sample per lane group,
// whole batch in one cluster of <= 8 CTAs): every input of a (sample, component)
// pair is requested up front -- ONE round trip to L2 -- and stays in registers
// through the forward, the backward and the eps-term fix-up; the two batch-wide sums
// (exp-sum for eps, the eps gradient) are cluster all-reduces through DSMEM.
// A sample is owned by PS*GW lanes: lane = (ph, k), k < GW components, and the P
// output dimensions are dealt round-robin over the PS "p-halves" (p = ii*PS + ph),
// which divides the per-lane transcendental work by PS; PL = ceil(P / PS) <= PLMAX.
#ifdef BSIG_NLL_PROF   // temporary instrumentation: %globaltimer marks of every CTA of the last launch
__device__ unsigned long long nll_prof_buf[8 * 16];
#define NLL_MARK(i)                                                        \
  if (threadIdx.x == 0) {                                                  \
    unsigned long long t_;                                                 \
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                 \
    nll_prof_buf[blockIdx.x * 16 + (i)] = t_;                              \
  }
#else
#define NLL_MARK(i)
#endif
template <int GW, int PS, int PLMAX, bool BWD>
__global__ void __launch_bounds__(512) nll_small_kernel(NllArgs a) {
  __shared__ float scratch[33];
  // batch-wide sums: every CTA PUSHES its partial into the inbox of every CTA of the cluster
  // (remote stores before a release/acquire cluster barrier), then adds the NC values from its
  // own shared memory in rank order -- no remote read latency, no barrier to keep a peer alive
  __shared__ float inbox[3][16];
  __shared__ float2 scratch2[33];
  cg::cluster_group cluster = cg::this_cluster();
  NLL_MARK(0)
  auto started = cluster.barrier_arrive();     // a peer's shared memory exists once it got here
  constexpr int LPS = GW * PS;             // lanes per sample (<= 32)
  const int tid = threadIdx.x;
  const int NC = gridDim.x, rank = blockIdx.x;
  const int SPB = blockDim.x / LPS;        // samples per CTA
  const int lane_s = tid & (LPS - 1), gid = tid / LPS;
  const int k = lane_s & (GW - 1), ph = lane_s / GW;
  const int B = a.B, P = a.P, K = a.K, PK = P * K;
  const int b = rank * SPB + gid;
  const bool ok = (b < B) && (k < K);
  const int64_t bb = ok ? b : 0;
  const int kk = ok ? k : 0;

  // ---- all loads (the minibatch row index is constant during a training call)
  const int64_t yrow = (a.y_rows ? __ldg(a.y_rows + bb) : bb) * P;
  pdl_wait_then_release();
  NLL_MARK(1)
  float zpi = ok ? __ldg(a.z_pi + bb * a.ld_pi + kk) : -INFINITY;
  float e[PLMAX], zi[PLMAX], nz[PLMAX];
#pragma unroll
  for (int ii = 0; ii < PLMAX; ++ii) {
    const int i = ii * PS + ph;
    const int ic = i < P ? i : 0;
    e[ii] = __ldg(a.zd + bb * a.ld_zd + ic * K + kk);
    zi[ii] = __ldg(a.y + yrow + ic) - __ldg(a.mu + bb * a.ld_mu + ic * K + kk);
    nz[ii] = __ldg(a.noise + bb * (int64_t)PK + ic * K + kk);
  }

  // ---- eps = 1e-5 * mean(exp(z_d)) over the batch
  float esum = 0.f;
  bool bad = false;
#pragma unroll
  for (int ii = 0; ii < PLMAX; ++ii) {
    e[ii] = xexp<true>(e[ii]);
    if (ok && ii * PS + ph < P) esum += e[ii];
  }
  NLL_MARK(2)
  esum = block_sum(esum, scratch);
  NLL_MARK(3)
  cluster.barrier_wait(std::move(started));
  if (tid < NC) *cluster.map_shared_rank(&inbox[0][rank], tid) = esum;
  auto pushed = cluster.barrier_arrive();
  NLL_MARK(4)

  // ---- mixture weights: softmax -> clamp -> renormalise (every p-half computes
  // the same values; xor offsets < GW stay inside one p-half); independent of eps, so it runs
  // while the partial exp-sums cross the cluster
  float mx = group_max<GW>(zpi);
  float soft = (zpi == -INFINITY) ? 0.f : xexp<true>(zpi - mx);
  const float sm = group_sum<GW>(soft);
  soft = soft / sm;
  float w = (k < K) ? fminf(fmaxf(soft, kMinWeight), 1.0f) : 0.f;
  const float csum = group_sum<GW>(w);
  w = w / csum;

  NLL_MARK(5)
  cluster.barrier_wait(std::move(pushed));
  NLL_MARK(6)
  float etot = 0.f;
  for (int r = 0; r < NC; ++r) etot += inbox[0][r];
  const float eps = kEpsNoise * (etot / (float)((int64_t)B * PK));

  // ---- log density of this component: partial sums over this lane's p, then
  // across the PS p-halves
  float quad = 0.f, logdet = 0.f;
#pragma unroll
  for (int ii = 0; ii < PLMAX; ++ii) {
    if (ii * PS + ph < P) {
      const float ldv = e[ii] + nz[ii] * eps;
      bad |= ok && !(finite_f(ldv) && finite_f(zi[ii]));
      zi[ii] = xdiv<true>(zi[ii], ldv);    // z = (y - mu) / L_d
      quad += zi[ii] * zi[ii];
      logdet += xlog<true>(ldv);
    }
  }
#pragma unroll
  for (int o = GW; o < LPS; o <<= 1) {
    quad += __shfl_xor_sync(0xffffffffu, quad, o);
    logdet += __shfl_xor_sync(0xffffffffu, logdet, o);
  }
  const float gj = -0.5f * ((float)P * kLog2Pi + quad) - logdet;
  bad |= ok && !(finite_f(gj) && finite_f(w));
  const float wc = fminf(fmaxf(w, kMinWeight), 1.0f);
  const float rk = ok ? fminf(fmaxf(gj, -kLLLimit), kLLLimit) + xlog<true>(wc) : -INFINITY;
  mx = group_max<GW>(rk);
  const float se = group_sum<GW>(rk == -INFINITY ? 0.f : xexp<true>(rk - mx));
  const float lse = mx + xlog<true>(se);
  float loss_acc = (b < B && lane_s == 0) ? -lse : 0.f;
  NLL_MARK(7)
  if (bad) atomicOr(a.flag, 1);

  float s_acc = 0.f;
  if (BWD) {
    const float coef = ok ? -xexp<true>(rk - lse) / (float)B : 0.f;
    const bool in_w = (w >= kMinWeight) && (w <= 1.0f);
    const float dw = (ok && in_w) ? coef / wc : 0.f;
    const float t1 = group_sum<GW>(dw * w);
    const float dc = (dw - t1) / csum;
    const float dp = ((soft >= kMinWeight) && (soft <= 1.0f)) ? dc : 0.f;
    const float t2 = group_sum<GW>(dp * soft);
    if (ok && ph == 0) a.d_pi[bb * a.ldo_pi + k] = soft * (dp - t2);
    const float cg = ((gj >= -kLLLimit) && (gj <= kLLLimit)) ? coef : 0.f;
#pragma unroll
    for (int ii = 0; ii < PLMAX; ++ii) {
      const int i = ii * PS + ph;
      if (i < P) {
        const float ldv = e[ii] + nz[ii] * eps;
        const float inv = xdiv<true>(1.0f, ldv);
        const float vi = zi[ii] * inv;
        const float dld = cg * (vi * zi[ii] - inv);
        if (ok) a.d_mu[bb * a.ldo_mu + i * K + k] = cg * vi;
        s_acc += ok ? dld * nz[ii] : 0.f;
        zi[ii] = dld;                      // keep d L_d for the final store
      }
    }
  }
  NLL_MARK(8)
  float lsum, ssum = 0.f;
  if (BWD) {               // both sums in one pass over the block
    const float2 ls = block_sum2(make_float2(loss_acc, s_acc), scratch2);
    lsum = ls.x; ssum = ls.y;
  } else {
    lsum = block_sum(loss_acc, scratch);
  }
  NLL_MARK(9)
  if (tid < NC) {
    *cluster.map_shared_rank(&inbox[1][rank], tid) = lsum;
    *cluster.map_shared_rank(&inbox[2][rank], tid) = ssum;
  }
  cluster.sync();
  NLL_MARK(10)
  if (rank == 0 && tid == 0) {
    float ltot = 0.f;
    for (int r = 0; r < NC; ++r) ltot += inbox[1][r];
    a.loss[0] = ltot / (float)B;
  }
  if (BWD) {
    float S = 0.f;
    for (int r = 0; r < NC; ++r) S += inbox[2][r];
    const float c = kEpsNoise * S / (float)((int64_t)B * PK);
    if (ok) {
#pragma unroll
      for (int ii = 0; ii < PLMAX; ++ii) {
        const int i = ii * PS + ph;
        if (i < P) a.d_zd[bb * a.ldo_zd + i * K + k] = e[ii] * (zi[ii] + c);
      }
    }
  }
  NLL_MARK(11)
}

// eps-term fix-up of the fused backward: dzd += exp(zd) * (1e-5/M) * S
__global__ void __launch_bounds__(256)
eps_fixup_kernel(const float* __restrict__ zd, int64_t ld_zd, float* dzd, int64_t ldo_zd,
                 int B, int PK, const float* s_parts, int nparts) {
  __shared__ float scratch[33];
  const float S = sum_parts(s_parts, nparts, scratch);
  const float c = kEpsNoise * S / (float)((int64_t)B * PK);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t start = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (((PK | ld_zd | ldo_zd) & 1) == 0 &&
      ((reinterpret_cast<uintptr_t>(zd) | reinterpret_cast<uintptr_t>(dzd)) & 7) == 0) {
    const int W = PK >> 1;
    const int64_t total = (int64_t)B * W;
    RowCol rc(start, stride, W);
    for (int64_t i = start; i < total; i += stride, rc.next()) {
      const float2 z2 = __ldg(reinterpret_cast<const float2*>(zd + rc.row * ld_zd) + rc.col);
      float2* dp = reinterpret_cast<float2*>(dzd + rc.row * ldo_zd) + rc.col;
      float2 d2 = *dp;
      d2.x += expf(z2.x) * c;
      d2.y += expf(z2.y) * c;
      *dp = d2;
    }
  } else {
    const int64_t total = (int64_t)B * PK;
    RowCol rc(start, stride, PK);
    for (int64_t i = start; i < total; i += stride, rc.next())
      dzd[rc.row * ldo_zd + rc.col] += expf(__ldg(zd + rc.row * ld_zd + rc.col)) * c;
  }
}

// --------------------------------------------------------- head epilogue (API path)
__global__ void __launch_bounds__(256)
head_fwd_kernel(const float* __restrict__ z, const float* __restrict__ noise,
                float* __restrict__ weights, float* __restrict__ l_d,
                int B, int P, int K, int NH, const float* ws, int nparts, int* flag) {
  __shared__ float scratch[33];
  const float esum = sum_parts(ws, nparts, scratch);
  const int PK = P * K;
  const float eps = kEpsNoise * (esum / (float)((int64_t)B * PK));
  bool bad = false;
  const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t gstride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t b = gtid; b < B; b += gstride) {
    const float* zr = z + b * NH;
    float mx = -INFINITY;
    for (int k = 0; k < K; ++k) mx = fmaxf(mx, __ldg(zr + k));
    float sm = 0.f;
    for (int k = 0; k < K; ++k) sm += expf(__ldg(zr + k) - mx);
    float cs = 0.f;
    for (int k = 0; k < K; ++k)
      cs += fminf(fmaxf(expf(__ldg(zr + k) - mx) / sm, kMinWeight), 1.0f);
    for (int k = 0; k < K; ++k) {
      const float wv = fminf(fmaxf(expf(__ldg(zr + k) - mx) / sm, kMinWeight), 1.0f) / cs;
      weights[b * K + k] = wv;
      bad |= !finite_f(wv);
    }
  }
  const int64_t total = (int64_t)B * PK;
  RowCol rc(gtid, gstride, PK);
  for (int64_t i = gtid; i < total; i += gstride, rc.next()) {
    const float v = expf(__ldg(z + rc.row * NH + K + PK + rc.col)) + __ldg(noise + i) * eps;
    l_d[i] = v;
    bad |= !finite_f(v) || !finite_f(__ldg(z + rc.row * NH + K + rc.col));
  }
  const int LK = NH - K - 2 * PK;
  if (LK > 0) {
    const int64_t total_l = (int64_t)B * LK;
    RowCol rl(gtid, gstride, LK);
    for (int64_t i = gtid; i < total_l; i += gstride, rl.next())
      bad |= !finite_f(__ldg(z + rl.row * NH + K + 2 * PK + rl.col));
  }
  if (bad) atomicOr(flag, 1);
}

// S = sum d_ld * noise (partials)
__global__ void __launch_bounds__(256)
dot_parts_kernel(const float* __restrict__ x, const float* __restrict__ y, int64_t total,
                 float* parts) {
  __shared__ float scratch[33];
  float acc = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x)
    acc += __ldg(x + i) * __ldg(y + i);
  acc = block_sum(acc, scratch);
  if (threadIdx.x == 0) parts[blockIdx.x] = acc;
}

__global__ void __launch_bounds__(256)
head_bwd_kernel(const float* __restrict__ z, const float* __restrict__ weights,
                const float* __restrict__ d_weights, const float* __restrict__ d_mu,
                const float* __restrict__ d_ld, const float* __restrict__ d_low,
                float* __restrict__ dz, int B, int P, int K, int NH,
                const float* s_parts, int nparts) {
  __shared__ float scratch[33];
  const float S = sum_parts(s_parts, nparts, scratch);
  const int PK = P * K;
  const float c_eps = kEpsNoise * S / (float)((int64_t)B * PK);
  const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t gstride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t b = gtid; b < B; b += gstride) {
    const float* zr = z + b * NH;
    float mx = -INFINITY;
    for (int k = 0; k < K; ++k) mx = fmaxf(mx, __ldg(zr + k));
    float sm = 0.f;
    for (int k = 0; k < K; ++k) sm += expf(__ldg(zr + k) - mx);
    float cs = 0.f, t1 = 0.f;
    for (int k = 0; k < K; ++k) {
      cs += fminf(fmaxf(expf(__ldg(zr + k) - mx) / sm, kMinWeight), 1.0f);
      t1 += __ldg(d_weights + b * K + k) * __ldg(weights + b * K + k);
    }
    float t2 = 0.f;
    for (int k = 0; k < K; ++k) {
      const float p = expf(__ldg(zr + k) - mx) / sm;
      const float dc = (__ldg(d_weights + b * K + k) - t1) / cs;
      const float dp = (p >= kMinWeight && p <= 1.0f) ? dc : 0.f;
      t2 += dp * p;
    }
    for (int k = 0; k < K; ++k) {
      const float p = expf(__ldg(zr + k) - mx) / sm;
      const float dc = (__ldg(d_weights + b * K + k) - t1) / cs;
      const float dp = (p >= kMinWeight && p <= 1.0f) ? dc : 0.f;
      dz[b * NH + k] = p * (dp - t2);
    }
  }
  const int64_t total = (int64_t)B * PK;
  RowCol rc(gtid, gstride, PK);
  for (int64_t i = gtid; i < total; i += gstride, rc.next()) {
    const int64_t o = rc.row * NH + K + rc.col;
    dz[o] = __ldg(d_mu + i);
    dz[o + PK] = expf(__ldg(z + o + PK)) * (__ldg(d_ld + i) + c_eps);
  }
  const int LK = NH - K - 2 * PK;
  if (LK > 0) {
    const int64_t total_l = (int64_t)B * LK;
    RowCol rl(gtid, gstride, LK);
    for (int64_t i = gtid; i < total_l; i += gstride, rl.next())
      dz[rl.row * NH + K + 2 * PK + rl.col] = d_low ? __ldg(d_low + i) : 0.f;
  }
}

// ------------------------------------------------------------------ host dispatch
template <int GW, int KPL, bool FUSED, bool FULL>
static int launch_nll_bwd_sel(const NllArgs& a, bool bwd, int grid, size_t smem, cudaStream_t st) {
  if (bwd) {
    if (smem > 48 * 1024)
      BSIG_CUDA(cudaFuncSetAttribute(nll_kernel<GW, KPL, FUSED, FULL, true>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    nll_kernel<GW, KPL, FUSED, FULL, true><<<grid, 128, smem, st>>>(a);
  } else {
    if (smem > 48 * 1024)
      BSIG_CUDA(cudaFuncSetAttribute(nll_kernel<GW, KPL, FUSED, FULL, false>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    nll_kernel<GW, KPL, FUSED, FULL, false><<<grid, 128, smem, st>>>(a);
  }
  BSIG_LAUNCH_CHECK();
  return 0;
}

template <int GW, int KPL>
static int launch_nll_gw(const NllArgs& a, bool fused, bool full, bool bwd, int grid, size_t smem,
                         cudaStream_t st) {
  if (fused) {
    return full ? launch_nll_bwd_sel<GW, KPL, true, true>(a, bwd, grid, smem, st)
                : launch_nll_bwd_sel<GW, KPL, true, false>(a, bwd, grid, smem, st);
  }
  return full ? launch_nll_bwd_sel<GW, KPL, false, true>(a, bwd, grid, smem, st)
              : launch_nll_bwd_sel<GW, KPL, false, false>(a, bwd, grid, smem, st);
}

static int launch_nll(NllArgs& a, bool fused, bool bwd, cudaStream_t st) {
  const int K = a.K;
  BSIG_REQUIRE(K >= 1 && K <= 128, "mixture NLL: 1 <= n_gaussians <= 128 supported (got %d)", K);
  BSIG_REQUIRE(a.P >= 1 && a.P <= 192, "mixture NLL: output_dim <= 192 supported (got %d)", a.P);
  const bool full = a.L > 0;
  int gw = 1;
  while (gw < K && gw < 32) gw <<= 1;
  const int gpb = 128 / gw;
  const int grid = (int)std::min<int64_t>(ceil_div(a.B, gpb), kMaxParts);
  const size_t smem = full ? (size_t)2 * a.P * 128 * sizeof(float) : 0;
  switch (gw) {
    case 1: return launch_nll_gw<1, 1>(a, fused, full, bwd, grid, smem, st);
    case 2: return launch_nll_gw<2, 1>(a, fused, full, bwd, grid, smem, st);
    case 4: return launch_nll_gw<4, 1>(a, fused, full, bwd, grid, smem, st);
    case 8: return launch_nll_gw<8, 1>(a, fused, full, bwd, grid, smem, st);
    case 16: return launch_nll_gw<16, 1>(a, fused, full, bwd, grid, smem, st);
    default:
      if (K <= 32) return launch_nll_gw<32, 1>(a, fused, full, bwd, grid, smem, st);
      if (K <= 64) return launch_nll_gw<32, 2>(a, fused, full, bwd, grid, smem, st);
      return launch_nll_gw<32, 4>(a, fused, full, bwd, grid, smem, st);
  }
}

template <int GW, int KPL, bool FULL, bool BWD>
static int launch_nll_cluster_t(const NllArgs& a, int nc, int tpb, size_t smem, cudaStream_t st) {
  if (smem > 48 * 1024)
    BSIG_CUDA(cudaFuncSetAttribute(nll_cluster_kernel<GW, KPL, FULL, BWD>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)nc);
  cfg.blockDim = dim3((unsigned)tpb);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)nc;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = add_pdl_attr(attr, 1);
  BSIG_CUDA(cudaLaunchKernelEx(&cfg, nll_cluster_kernel<GW, KPL, FULL, BWD>, a));
  BSIG_LAUNCH_CHECK();
  return 0;
}

template <int GW, int PS, int PLMAX, bool BWD>
static int launch_nll_small_t(const NllArgs& a, int nc, int tpb, cudaStream_t st) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)nc);
  cfg.blockDim = dim3((unsigned)tpb);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)nc;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = add_pdl_attr(attr, 1);
  if (nc > 8)
    BSIG_CUDA(cudaFuncSetAttribute(nll_small_kernel<GW, PS, PLMAX, BWD>,
                                   cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  BSIG_CUDA(cudaLaunchKernelEx(&cfg, nll_small_kernel<GW, PS, PLMAX, BWD>, a));
  BSIG_LAUNCH_CHECK();
  return 0;
}

template <int GW, int PS>
static int launch_nll_small_p(const NllArgs& a, bool bwd, int nc, int tpb, cudaStream_t st) {
  const int pl = (a.P + PS - 1) / PS;
#define BSIG_PL(PLV)                                                                    \
  if (pl <= PLV)                                                                        \
    return bwd ? launch_nll_small_t<GW, PS, PLV, true>(a, nc, tpb, st)                  \
               : launch_nll_small_t<GW, PS, PLV, false>(a, nc, tpb, st);
  BSIG_PL(4) BSIG_PL(8) BSIG_PL(16) BSIG_PL(40)
#undef BSIG_PL
  return -1;
}

// Register-resident minibatch form; -1 if not applicable.
static int launch_nll_small(const NllArgs& a, bool bwd, cudaStream_t st) {
  if (a.L > 0 || a.K > 32 || a.P > 40) return -1;
  int gw = 1;
  while (gw < a.K) gw <<= 1;
  // split the P dimension over up to 4 lanes per component while a sample still
  // fits one warp and a lane keeps at least two output dimensions
  int ps = 1;
  while (ps < 4 && gw * ps * 2 <= 32 && a.P >= 4 * ps) ps <<= 1;
  const int lps = gw * ps;
  // the kernel is issue-bound per SM (16 warps of dependent chains): 256-thread CTAs over a
  // cluster of up to 16 (non-portable size) halve every compute phase of the 100-row minibatch
  int tpb = 256;
  if (ceil_div(a.B, tpb / lps) > 16) tpb = 512;
  const int nc = (int)ceil_div(a.B, tpb / lps);
  if (nc > (tpb == 256 ? 16 : 8)) return -1;
#define BSIG_NS(GWV, PSV) \
  if (gw == GWV && ps == PSV) return launch_nll_small_p<GWV, PSV>(a, bwd, nc, tpb, st);
  BSIG_NS(1, 1) BSIG_NS(1, 2) BSIG_NS(1, 4) BSIG_NS(2, 1) BSIG_NS(2, 2) BSIG_NS(2, 4)
  BSIG_NS(4, 1) BSIG_NS(4, 2) BSIG_NS(4, 4) BSIG_NS(8, 1) BSIG_NS(8, 2) BSIG_NS(8, 4)
  BSIG_NS(16, 1) BSIG_NS(16, 2) BSIG_NS(32, 1)
#undef BSIG_NS
  return -1;
}

// Returns 0 if launched, -1 if the problem does not fit one cluster, >0 on error.
static int launch_nll_cluster(const NllArgs& a, bool bwd, cudaStream_t st) {
  const int K = a.K;
  if (K > 32) return -1;
  int gw = 1;
  while (gw < K) gw <<= 1;
  const int64_t lanes = (int64_t)a.B * gw;
  if (lanes > 8 * 512) return -1;
  const int tpb = lanes <= 8 * 128 ? 128 : (lanes <= 8 * 256 ? 256 : 512);
  const int nc = (int)ceil_div(lanes, tpb);
  const bool full = a.L > 0;
  const size_t smem = full ? (size_t)2 * a.P * tpb * sizeof(float) : 0;
  if (smem > 160 * 1024) return -1;
#define BSIG_CL(GWV)                                                                        \
  case GWV:                                                                                 \
    if (full) return bwd ? launch_nll_cluster_t<GWV, 1, true, true>(a, nc, tpb, smem, st)   \
                         : launch_nll_cluster_t<GWV, 1, true, false>(a, nc, tpb, smem, st); \
    return bwd ? launch_nll_cluster_t<GWV, 1, false, true>(a, nc, tpb, smem, st)            \
               : launch_nll_cluster_t<GWV, 1, false, false>(a, nc, tpb, smem, st);
  switch (gw) {
    BSIG_CL(1) BSIG_CL(2) BSIG_CL(4) BSIG_CL(8) BSIG_CL(16) BSIG_CL(32)
  }
#undef BSIG_CL
  return -1;
}

template <int GW, int KPL, bool FULL, bool BWD>
static int launch_nll_stream_t(const NllArgs& a, const NllStream& q, int threads, size_t smem,
                               int* nparts, cudaStream_t st) {
  auto kern = nll_stream_kernel<GW, KPL, FULL, BWD>;
  if (smem > 48 * 1024)
    BSIG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 1;
  BSIG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem));
  const int64_t ntiles = ceil_div(a.B, q.R);
  const int grid = (int)std::min<int64_t>(
      std::min<int64_t>(ntiles, (int64_t)std::max(occ, 1) * sm_count()), kMaxParts);
  kern<<<grid, threads, smem, st>>>(a, q);
  BSIG_LAUNCH_CHECK();
  *nparts = grid;
  return 0;
}

template <int GW, int KPL>
static int launch_nll_stream_gw(const NllArgs& a, const NllStream& q, bool full, bool bwd,
                                int threads, size_t smem, int* nparts, cudaStream_t st) {
  if (full)
    return bwd ? launch_nll_stream_t<GW, KPL, true, true>(a, q, threads, smem, nparts, st)
               : launch_nll_stream_t<GW, KPL, true, false>(a, q, threads, smem, nparts, st);
  return bwd ? launch_nll_stream_t<GW, KPL, false, true>(a, q, threads, smem, nparts, st)
             : launch_nll_stream_t<GW, KPL, false, false>(a, q, threads, smem, nparts, st);
}

// Streaming (bulk-copy staged) form of the fused NLL for batches beyond one cluster.
// Returns 0 if launched (*nparts = number of S partials), -1 if not applicable.
static int launch_nll_stream(const NllArgs& a, const float* z, float* dz, int64_t NH, bool bwd,
                             int* nparts, cudaStream_t st) {
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  if (!al16(z) || !al16(a.noise) || (bwd && !al16(dz))) return -1;
  const int K = a.K, P = a.P;
  if (K < 1 || K > 128 || P < 1 || P > 192) return -1;
  const bool full = a.L > 0;
  // K <= 32: one component per lane, groups of exactly K lanes packed into the warps
  const int rpw = K <= 32 ? 32 / K : 1;                     // samples per warp and pass
  const int64_t PK = (int64_t)P * K;
  const int64_t per_row = 2 * (NH + PK + P + ((bwd && full) ? NH : 0)) * 4;
  // tile height R (multiple of 4) and CTA width (consumer warps + the producer warp):
  // maximise (samples resident on an SM) x (fraction of lane groups that own a sample)
  int best_r = 0, best_threads = 0;
  double best_score = 0.0;
  for (int r = 4; r <= 8 * rpw && r <= 64; r += 4) {
    const int warps = (int)ceil_div(r, rpw);
    if (warps > 8) break;
    const int64_t smem = r * per_row + (full ? 2 * (int64_t)P * warps * 32 * 4 : 0) + 64;
    if (smem > 220 * 1024) break;
    const int ctas = (int)std::min<int64_t>(
        std::min<int64_t>((226 * 1024) / (smem + 1024), 64 / (warps + 1)), 32);
    const double score = (double)ctas * r * ((double)r / (warps * rpw));
    if (score >= best_score) { best_score = score; best_r = r; best_threads = (warps + 1) * 32; }
  }
  if (best_r < 4) return -1;
  NllStream q;
  q.z = z; q.dz = dz; q.NH = (int)NH; q.R = best_r;
  const size_t smem =
      (size_t)best_r * per_row + (full ? (size_t)2 * P * (best_threads - 32) * 4 : 0) + 64;
  if (K <= 32) return launch_nll_stream_gw<0, 1>(a, q, full, bwd, best_threads, smem, nparts, st);
  if (K <= 64) return launch_nll_stream_gw<32, 2>(a, q, full, bwd, best_threads, smem, nparts, st);
  return launch_nll_stream_gw<32, 4>(a, q, full, bwd, best_threads, smem, nparts, st);
}

static int launch_exp_sum(const float* zd, int64_t ld_zd, int B, int PK, float* ws,
                          cudaStream_t st, int* nparts) {
  const int64_t total = (int64_t)B * PK;
  const int grid = (int)std::min<int64_t>(ceil_div(total, 256 * 4), kMaxParts);
  exp_sum_kernel<<<grid, 256, 0, st>>>(zd, ld_zd, B, PK, ws);
  BSIG_LAUNCH_CHECK();
  *nparts = grid;
  return 0;
}

}  // namespace bsig

using namespace bsig;

extern "C" int64_t bsig_mdn_ws_bytes(int64_t b) {
  (void)b;
  return (int64_t)kWsFloats * sizeof(float);
}

extern "C" int bsig_mdn_head_fwd(const float* z, const float* noise, float* weights, float* l_d,
                                 int64_t b, int64_t p, int64_t k, int full_cov, void* ws,
                                 int64_t ws_bytes, int* flag, void* stream) {
  BSIG_REQUIRE(ws_bytes >= bsig_mdn_ws_bytes(b), "mdn_head_fwd: workspace too small");
  BSIG_REQUIRE(b >= 1 && p >= 1 && k >= 1, "mdn_head_fwd: bad sizes");
  const int L = (full_cov && p > 1) ? (int)(p * (p - 1) / 2) : 0;
  const int NH = (int)(k + 2 * p * k + (int64_t)L * k);
  cudaStream_t st = (cudaStream_t)stream;
  int nparts = 0;
  if (launch_exp_sum(z + k + p * k, NH, (int)b, (int)(p * k), (float*)ws, st, &nparts)) return 1;
  const int64_t work = b * (int64_t)NH;
  const int grid = (int)std::min<int64_t>(ceil_div(work, 256), (int64_t)sm_count() * 8);
  head_fwd_kernel<<<grid, 256, 0, st>>>(z, noise, weights, l_d, (int)b, (int)p, (int)k, NH,
                                        (const float*)ws, nparts, flag);
  BSIG_LAUNCH_CHECK();
  return 0;
}

extern "C" int bsig_mdn_head_bwd(const float* z, const float* noise, const float* weights,
                                 const float* d_weights, const float* d_mu, const float* d_ld,
                                 const float* d_low, float* dz, int64_t b, int64_t p, int64_t k,
                                 int full_cov, void* ws, int64_t ws_bytes, void* stream) {
  BSIG_REQUIRE(ws_bytes >= bsig_mdn_ws_bytes(b), "mdn_head_bwd: workspace too small");
  const int L = (full_cov && p > 1) ? (int)(p * (p - 1) / 2) : 0;
  const int NH = (int)(k + 2 * p * k + (int64_t)L * k);
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t total = b * p * k;
  float* s_parts = (float*)ws + 2 * kMaxParts;
  const int gparts = (int)std::min<int64_t>(ceil_div(total, 1024), 256);
  dot_parts_kernel<<<gparts, 256, 0, st>>>(d_ld, noise, total, s_parts);
  BSIG_LAUNCH_CHECK();
  const int grid = (int)std::min<int64_t>(ceil_div(b * (int64_t)NH, 256), (int64_t)sm_count() * 8);
  head_bwd_kernel<<<grid, 256, 0, st>>>(z, weights, d_weights, d_mu, d_ld, d_low, dz, (int)b,
                                        (int)p, (int)k, NH, s_parts, gparts);
  BSIG_LAUNCH_CHECK();
  return 0;
}

static void fill_common(NllArgs& a, const float* y, const int64_t* y_rows, float* loss, void* ws,
                        int* flag, int64_t b, int64_t p, int64_t k, int L) {
  a.y = y; a.y_rows = y_rows; a.loss = loss; a.ws = (float*)ws; a.flag = flag;
  a.B = (int)b; a.P = (int)p; a.K = (int)k; a.L = L;
  a.grad_scale = nullptr; a.noise = nullptr; a.nparts_e = 0;
  a.d_pi = a.d_mu = a.d_zd = a.d_low = nullptr;
  a.ldo_pi = a.ldo_mu = a.ldo_zd = a.ldo_low = 0;
}

extern "C" int bsig_mog_nll_fwd(const float* weights, const float* mu, int64_t ld_mu,
                                const float* l_d, int64_t ld_ld, const float* low, int64_t ld_low,
                                const float* y, const int64_t* y_rows, float* loss, int64_t b,
                                int64_t p, int64_t k, void* ws, int64_t ws_bytes, int* flag,
                                void* stream) {
  BSIG_REQUIRE(ws_bytes >= bsig_mdn_ws_bytes(b), "mog_nll_fwd: workspace too small");
  BSIG_REQUIRE(b >= 1, "mog_nll_fwd: empty batch");
  NllArgs a;
  fill_common(a, y, y_rows, loss, ws, flag, b, p, k, low ? (int)(p * (p - 1) / 2) : 0);
  a.z_pi = weights; a.ld_pi = k; a.mu = mu; a.ld_mu = ld_mu; a.zd = l_d; a.ld_zd = ld_ld;
  a.low = low; a.ld_low = ld_low;
  return launch_nll(a, false, false, (cudaStream_t)stream);
}

extern "C" int bsig_mog_nll_bwd(const float* weights, const float* mu, int64_t ld_mu,
                                const float* l_d, int64_t ld_ld, const float* low, int64_t ld_low,
                                const float* y, const int64_t* y_rows, const float* grad_scale,
                                float* d_weights, float* d_mu, float* d_ld, float* d_low,
                                int64_t b, int64_t p, int64_t k, void* ws, int64_t ws_bytes,
                                void* stream) {
  BSIG_REQUIRE(ws_bytes >= bsig_mdn_ws_bytes(b), "mog_nll_bwd: workspace too small");
  BSIG_REQUIRE(b >= 1, "mog_nll_bwd: empty batch");
  NllArgs a;
  // the (re-computed) loss and the finite flag of the backward launch are
  // discarded into two spare workspace slots
  float* wsf = (float*)ws;
  fill_common(a, y, y_rows, wsf + 3 * kMaxParts + 4, ws, (int*)(wsf + 3 * kMaxParts + 5), b, p,
              k, low ? (int)(p * (p - 1) / 2) : 0);
  a.z_pi = weights; a.ld_pi = k; a.mu = mu; a.ld_mu = ld_mu; a.zd = l_d; a.ld_zd = ld_ld;
  a.low = low; a.ld_low = ld_low; a.grad_scale = grad_scale;
  a.d_pi = d_weights; a.ldo_pi = k;
  a.d_mu = d_mu; a.ldo_mu = p * k;
  a.d_zd = d_ld; a.ldo_zd = p * k;
  a.d_low = d_low; a.ldo_low = low ? (p * (p - 1) / 2) * k : 0;
  return launch_nll(a, false, true, (cudaStream_t)stream);
}

extern "C" int bsig_mdn_nll_fused(const float* z, const float* noise, const float* y,
                                  const int64_t* y_rows, float* loss, float* dz, int64_t b,
                                  int64_t p, int64_t k, int full_cov, void* ws, int64_t ws_bytes,
                                  int* flag, void* stream) {
  BSIG_REQUIRE(ws_bytes >= bsig_mdn_ws_bytes(b), "mdn_nll_fused: workspace too small");
  BSIG_REQUIRE(b >= 1 && p >= 1 && k >= 1, "mdn_nll_fused: bad sizes");
  const int L = (full_cov && p > 1) ? (int)(p * (p - 1) / 2) : 0;
  const int64_t PK = p * k;
  const int64_t NH = k + 2 * PK + (int64_t)L * k;
  cudaStream_t st = (cudaStream_t)stream;
  NllArgs a;
  fill_common(a, y, y_rows, loss, ws, flag, b, p, k, L);
  a.z_pi = z; a.ld_pi = NH;
  a.mu = z + k; a.ld_mu = NH;
  a.zd = z + k + PK; a.ld_zd = NH;
  a.low = L ? z + k + 2 * PK : nullptr; a.ld_low = NH;
  a.noise = noise;
  const bool bwd = dz != nullptr;
  if (bwd) {
    a.d_pi = dz; a.ldo_pi = NH;
    a.d_mu = dz + k; a.ldo_mu = NH;
    a.d_zd = dz + k + PK; a.ldo_zd = NH;
    a.d_low = L ? dz + k + 2 * PK : nullptr; a.ldo_low = NH;
  }
  {
    int rc = launch_nll_small(a, bwd, st);           // register-resident minibatch form
    if (rc >= 0) return rc;
    rc = launch_nll_cluster(a, bwd, st);             // one launch when the batch fits a cluster
    if (rc >= 0) return rc;
  }
  if (launch_exp_sum(a.zd, NH, (int)b, (int)PK, (float*)ws, st, &a.nparts_e)) return 1;
  int nparts = 0;
  const bool no_stream = getenv("BSIG_NLL_NO_STREAM") != nullptr;     // A/B switch (profiling)
  int rc = no_stream ? -1 : launch_nll_stream(a, z, dz, NH, bwd, &nparts, st);
  if (rc > 0) return rc;
  if (rc < 0) {
    if (launch_nll(a, true, bwd, st)) return 1;
    int gw = 1;
    while (gw < k && gw < 32) gw <<= 1;
    nparts = (int)std::min<int64_t>(ceil_div(b, 128 / gw), kMaxParts);
  }
  if (bwd) {
    const int grid = (int)std::min<int64_t>(ceil_div(b * PK, 256), (int64_t)sm_count() * 8);
    eps_fixup_kernel<<<grid, 256, 0, st>>>(a.zd, NH, a.d_zd, NH, (int)b, (int)PK,
                                           (float*)ws + 2 * kMaxParts, nparts);
    BSIG_LAUNCH_CHECK();
  }
  return 0;
}

#ifdef BSIG_NLL_PROF
extern "C" int dbg_nll_prof_read(unsigned long long* out) {
  return (int)cudaMemcpyFromSymbol(out, bsig::nll_prof_buf, sizeof(unsigned long long) * 8 * 16);
}
#endif
