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
#include <algorithm>
#include <cstdlib>

#include "async_copy.cuh"
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
  int bulk;         // small-C kernel: tiles may travel by cp.async.bulk
  int raw_stride;   // small-C kernel: floats per raw-rollout buffer
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

// Increments d[traj][t][0..CPAD) (zero beyond C) in shared memory; rows of CPAD
// floats are 16B aligned so that a step's increment is read with LDS.128.
// Two passes: (1) coalesced, independent loads of the raw rollouts into `raw`
// (memory-level parallelism: nothing depends on a previous load), (2) the
// differences from shared memory.  `raw` may alias the output staging buffer.
template <int CPAD>
__device__ __forceinline__ void load_increments(const SigArgs& p, float* ds, int traj_stride,
                                                float* raw, int64_t traj0, int ntraj) {
  const int LD = p.L * p.D, LA = p.L * p.A;
  float* raw_s = raw;                       // [ntraj][L*D]
  float* raw_a = raw + (size_t)p.tpb * LD;  // [ntraj][L*A]
  const float* gs = p.states + traj0 * p.s_stride;
  const float* ga = p.actions + traj0 * p.a_stride;
  if (p.s_stride == LD) {
#pragma unroll 4
    for (int e = threadIdx.x; e < ntraj * LD; e += blockDim.x) raw_s[e] = __ldg(gs + e);
  } else {
#pragma unroll 4
    for (int e = threadIdx.x; e < ntraj * LD; e += blockDim.x) {
      const int tl = e / LD, r = e - tl * LD;
      raw_s[e] = __ldg(gs + tl * p.s_stride + r);
    }
  }
  if (p.a_stride == LA) {
#pragma unroll 4
    for (int e = threadIdx.x; e < ntraj * LA; e += blockDim.x) raw_a[e] = __ldg(ga + e);
  } else {
#pragma unroll 4
    for (int e = threadIdx.x; e < ntraj * LA; e += blockDim.x) {
      const int tl = e / LA, r = e - tl * LA;
      raw_a[e] = __ldg(ga + tl * p.a_stride + r);
    }
  }
  __syncthreads();
  const int steps = p.L - 1;
  for (int e = threadIdx.x; e < ntraj * steps; e += blockDim.x) {
    const int tl = e / steps, t = e - tl * steps;
    float* drow = ds + tl * traj_stride + t * CPAD;
    const float* s = raw_s + tl * LD + t * p.D;
    const float* a = raw_a + tl * LA + t * p.A;
    drow[0] = 1.0f;                                       // time channel t+1 -> t+2
    for (int c = 0; c < p.D; ++c) drow[1 + c] = s[p.D + c] - s[c];
    for (int c = 0; c < p.A; ++c) drow[1 + p.D + c] = a[p.A + c] - a[c];
    for (int c = p.C; c < CPAD; ++c) drow[c] = 0.f;
  }
}

__device__ __forceinline__ void store_staged(const SigArgs& p, const float* stage, int64_t traj0,
                                             int ntraj) {
  float* out = p.out + traj0 * p.siglen;
  const int64_t total = (int64_t)ntraj * p.siglen;
  if ((reinterpret_cast<uintptr_t>(out) & 15) == 0) {
    const int64_t n4 = total >> 2;
    for (int64_t e = threadIdx.x; e < n4; e += blockDim.x)
      st_stream_f4(out + 4 * e, *reinterpret_cast<const float4*>(stage + 4 * e));
    for (int64_t e = 4 * n4 + threadIdx.x; e < total; e += blockDim.x) out[e] = stage[e];
  } else {
    for (int64_t e = threadIdx.x; e < total; e += blockDim.x) out[e] = stage[e];
  }
}

// Small channel counts (C <= 8; Pendulum C=5, Cartpole C=6): one thread owns row i
// of every level -- S1[i], S2[i,:], S3[i,:,:] (C*C registers) -- so a step is
// C*C + 2C FFMAs.  Persistent, warp-specialised CTAs walk tiles of `tpb` trajectories
// (one producer warp, the others consumers, synchronised through mbarriers only):
//   * the raw rollouts of a tile are contiguous in HBM ([tpb][L*D], [tpb][L*A]) and
//     arrive by cp.async.bulk into a double buffer (tile it+2 is in flight while
//     tile it is being computed) -- no load instructions, no exposed latency;
//   * increments are formed on the fly from the raw rows (per-channel shared-memory
//     cursors), so there is no separate differencing pass;
//   * the finished levels are staged densely in shared memory and leave as ONE bulk
//     store per tile, which overlaps the next tile's recursion.
// Tiles that cannot use bulk copies (strided rollouts, misaligned pointers, a ragged
// last tile) take plain cooperative loads / stores through the same buffers.
__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}

template <int C, int JP>      // JP: level-3 rows updated with packed FFMA2 (the rest: FFMA)
__global__ void __launch_bounds__(256, (C >= 7 ? 2 : 3)) signature3_small_kernel(SigArgs p) {
  extern __shared__ __align__(16) float smem[];
  // full[b]: raw buffer b holds a tile;  done: the consumers have finished a tile (raw
  // buffer free, stage written);  stage_free: the previous tile's bulk store has read
  // the stage
  __shared__ __align__(8) uint64_t full[2], done, stage_free;
  const int tid = threadIdx.x, lane = tid & 31;
  const int TC = blockDim.x - 32;                        // consumer threads; last warp = producer
  const bool producer = tid >= TC;
  const int L = p.L, D = p.D, A = p.A, LD = L * D, LA = L * A;
  const int steps = L - 1, tpb = p.tpb;
  constexpr int CP = (C + 1) / 2;                          // packed pairs per level-3 row
  auto rawb = [&](int b) { return smem + (size_t)b * p.raw_stride; };   // each [tpb][L*D] | [tpb][L*A]
  float* stage = smem + 2 * (size_t)p.raw_stride;        // [tpb][siglen]
  const int64_t ntiles = (p.n + tpb - 1) / tpb;
  const int my_tiles = (int64_t)blockIdx.x < ntiles
                           ? (int)((ntiles - 1 - blockIdx.x) / gridDim.x) + 1 : 0;
  auto tile_of = [&](int it) { return (int64_t)blockIdx.x + (int64_t)it * gridDim.x; };
  auto tile_rows = [&](int64_t tile) { return (int)min((int64_t)tpb, p.n - tile * tpb); };
  auto tile_bulk = [&](int64_t tile) { return p.bulk && (tile_rows(tile) & 3) == 0; };

  if (tid == 0) {
    ac::mbar_init(&full[0], 1);
    ac::mbar_init(&full[1], 1);
    ac::mbar_init(&done, TC >> 5);
    ac::mbar_init(&stage_free, 1);
    ac::fence_barrier_init();
  }
  __syncthreads();

  if (producer) {
    // ---- producer warp: loads run two tiles ahead, stores trail the consumers
    auto load_tile = [&](int it) {
      const int b = it & 1;
      const int64_t tile = tile_of(it), traj0 = tile * tpb;
      const int ntraj = tile_rows(tile);
      float* raw_s = rawb(b);
      float* raw_a = rawb(b) + (size_t)tpb * LD;
      if (tile_bulk(tile)) {
        if (lane == 0) {
          const uint32_t bs = (uint32_t)ntraj * LD * 4, ba = (uint32_t)ntraj * LA * 4;
          ac::mbar_expect_tx(&full[b], bs + ba);
          ac::bulk_g2s(raw_s, p.states + traj0 * LD, bs, &full[b]);
          if (ba) ac::bulk_g2s(raw_a, p.actions + traj0 * LA, ba, &full[b]);
        }
      } else {
        for (int e = lane; e < ntraj * LD; e += 32) {
          const int r = e / LD, c = e - r * LD;
          raw_s[e] = __ldg(p.states + (traj0 + r) * p.s_stride + c);
        }
        for (int e = lane; e < ntraj * LA; e += 32) {
          const int r = e / LA, c = e - r * LA;
          raw_a[e] = __ldg(p.actions + (traj0 + r) * p.a_stride + c);
        }
        __syncwarp();
        if (lane == 0) ac::mbar_arrive(&full[b]);
      }
    };
    if (my_tiles > 0) load_tile(0);
    if (my_tiles > 1) load_tile(1);
    for (int it = 0; it < my_tiles; ++it) {
      const int64_t tile = tile_of(it), traj0 = tile * tpb;
      const int ntraj = tile_rows(tile);
      ac::mbar_wait(&done, (uint32_t)it & 1u);           // tile computed and staged
      if (tile_bulk(tile)) {
        if (lane == 0) {
          ac::bulk_s2g(p.out + traj0 * p.siglen, stage, (uint32_t)ntraj * (uint32_t)p.siglen * 4u);
          ac::bulk_commit();
        }
      } else {
        float* out = p.out + traj0 * p.siglen;
        const int64_t total = (int64_t)ntraj * p.siglen;
        for (int64_t e = lane; e < total; e += 32) out[e] = stage[e];
      }
      if (it + 2 < my_tiles) load_tile(it + 2);          // raw buffer it & 1 is free again
      if (lane == 0) ac::bulk_wait_read<0>();
      __syncwarp();
      if (lane == 0) ac::mbar_arrive(&stage_free);
    }
    return;
  }

  // ---- consumer warps
  const int tl = tid / C, i = tid - tl * C;
  for (int it = 0; it < my_tiles; ++it) {
    const int b = it & 1;
    const int ntraj = tile_rows(tile_of(it));
    ac::mbar_wait(&full[b], (uint32_t)(it >> 1) & 1u);

    // level-3 rows are kept as packed pairs (k, k+1) and updated with FFMA2 (two fp32
    // FMAs per issued instruction, same rounding as fmaf); odd C pads each row by one
    float s1 = 0.f, s2[C];
    float2 s3[C][CP];
#pragma unroll
    for (int j = 0; j < C; ++j) {
      s2[j] = 0.f;
#pragma unroll
      for (int k = 0; k < CP; ++k) s3[j][k] = make_float2(0.f, 0.f);
    }
    const bool active = tl < ntraj;
    if (active) {
      // shared-memory cursor (32-bit shared address) of every channel; channel 0 is
      // the time channel, whose increment is exactly 1
      const uint32_t base = ac::smem_u32(rawb(b));
      const uint32_t s_off = base + 4u * (uint32_t)(tl * LD);
      const uint32_t a_off = base + 4u * (uint32_t)(tpb * LD + tl * LA);
      // all state channels share one cursor, all action channels another (shifted by
      // -4D so that channel c sits at offset 4(c-1) from either)
      uint32_t s_cur = s_off, a_cur = a_off - 4u * (uint32_t)D;
      const uint32_t s_str = 4u * (uint32_t)D, a_str = 4u * (uint32_t)A;
      float prev[C];
#pragma unroll
      for (int c = 1; c < C; ++c) prev[c] = lds_f32((c <= D ? s_cur : a_cur) + 4u * (c - 1));
      const bool own_s = i <= D;
      uint32_t ocur = (i == 0) ? s_off : (own_s ? s_off + 4u * (i - 1) : a_off + 4u * (i - 1 - D));
      const uint32_t ostr = (i == 0) ? 0u : (own_s ? s_str : a_str);
      float oprev = lds_f32(ocur);
#pragma unroll 2
      for (int t = 0; t < steps; ++t) {
        float dv[2 * CP];
        dv[0] = 1.0f;
        dv[2 * CP - 1] = 0.f;                 // pad lane of odd C
        s_cur += s_str;
        a_cur += a_str;
#pragma unroll
        for (int c = 1; c < C; ++c) {
          const float x = lds_f32((c <= D ? s_cur : a_cur) + 4u * (c - 1));
          dv[c] = x - prev[c];
          prev[c] = x;
        }
        ocur += ostr;
        const float ox = lds_f32(ocur);
        const float di = (i == 0) ? 1.0f : ox - oprev;
        oprev = ox;
        const float a3 = (s1 + di * (1.0f / 3.0f)) * 0.5f;
        const float a2 = s1 + di * 0.5f;
#pragma unroll
        for (int j = 0; j < C; ++j) {
          const float t2 = fmaf(a3, dv[j], s2[j]);
          if (j < JP) {                      // packed rows: FFMA2 (fma-heavy pipe only)
            const float2 t22 = make_float2(t2, t2);
#pragma unroll
            for (int k = 0; k < CP; ++k)
              s3[j][k] = __ffma2_rn(t22, make_float2(dv[2 * k], dv[2 * k + 1]), s3[j][k]);
          } else {                           // scalar rows: FFMA (either fma pipe)
#pragma unroll
            for (int k = 0; k < CP; ++k) {
              s3[j][k].x = fmaf(t2, dv[2 * k], s3[j][k].x);
              s3[j][k].y = fmaf(t2, dv[2 * k + 1], s3[j][k].y);
            }
          }
          s2[j] = fmaf(a2, dv[j], s2[j]);
        }
        s1 += di;
      }
    }
    // the previous tile's bulk store must have finished reading `stage`
    if (it > 0) ac::mbar_wait(&stage_free, (uint32_t)(it - 1) & 1u);
    if (active) {
      float* o = stage + (size_t)tl * p.siglen;
      o[i] = s1;
      if ((C & 1) == 0) {                  // even C: every block below starts at an even offset
        float2* o2 = reinterpret_cast<float2*>(o + C + i * C);
#pragma unroll
        for (int j = 0; j < C / 2; ++j) o2[j] = make_float2(s2[2 * j], s2[2 * j + 1]);
        float2* o3 = reinterpret_cast<float2*>(o + C + C * C + i * C * C);
#pragma unroll
        for (int j = 0; j < C; ++j)
#pragma unroll
          for (int k = 0; k < CP; ++k) o3[j * CP + k] = s3[j][k];
      } else {
#pragma unroll
        for (int j = 0; j < C; ++j) {
          o[C + i * C + j] = s2[j];
#pragma unroll
          for (int k = 0; k < C; ++k)
            o[C + C * C + (i * C + j) * C + k] = (k & 1) ? s3[j][k / 2].y : s3[j][k / 2].x;
        }
      }
    }
    ac::fence_proxy_async();               // staged signature -> visible to the bulk store
    __syncwarp();
    if (lane == 0) ac::mbar_arrive(&done);
  }
}

// 9 <= C <= 22 (depth 3 stops at C = 22): one thread owns one (i,j) pair --
// S2[i,j] and the vector S3[i,j,:] in registers.
template <int CPAD>
__global__ void __launch_bounds__(512) signature3_kernel(SigArgs p) {
  extern __shared__ __align__(16) float smem[];
  const int C = p.C, CC = C * C;
  const int steps = p.L - 1;
  const int tstride = steps * CPAD;
  float* ds = smem;                                     // [tpb][steps][CPAD]
  float* stage = smem + (size_t)p.tpb * tstride;        // [tpb][siglen]
  const int64_t traj0 = (int64_t)blockIdx.x * p.tpb;
  const int ntraj = (int)min((int64_t)p.tpb, p.n - traj0);
  load_increments<CPAD>(p, ds, tstride, stage, traj0, ntraj);
  __syncthreads();

  const int tl = threadIdx.x / CC;
  const int ij = threadIdx.x - tl * CC;
  const int i = ij / C, j = ij - i * C;
  if (tl < ntraj) {
    const float* d = ds + tl * tstride;
    float s1 = 0.f, s2 = 0.f;
    float s3[CPAD];
#pragma unroll
    for (int k = 0; k < CPAD; ++k) s3[k] = 0.f;
    for (int t = 0; t < steps; ++t) {
      const float* dt = d + t * CPAD;
      const float di = dt[i], dj = dt[j];
      const float t2 = fmaf((s1 + di * (1.0f / 3.0f)) * 0.5f, dj, s2);
#pragma unroll
      for (int k4 = 0; k4 < CPAD / 4; ++k4) {
        const float4 dk = *reinterpret_cast<const float4*>(dt + 4 * k4);
        s3[4 * k4 + 0] = fmaf(t2, dk.x, s3[4 * k4 + 0]);
        s3[4 * k4 + 1] = fmaf(t2, dk.y, s3[4 * k4 + 1]);
        s3[4 * k4 + 2] = fmaf(t2, dk.z, s3[4 * k4 + 2]);
        s3[4 * k4 + 3] = fmaf(t2, dk.w, s3[4 * k4 + 3]);
      }
      s2 = fmaf(s1 + di * 0.5f, dj, s2);
      s1 += di;
    }
    float* o = stage + (size_t)tl * p.siglen;
    if (j == 0) o[i] = s1;
    o[C + ij] = s2;
#pragma unroll
    for (int k = 0; k < CPAD; ++k)
      if (k < C) o[C + CC + ij * C + k] = s3[k];
  }
  __syncthreads();
  store_staged(p, stage, traj0, ntraj);
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
    const int64_t steps = len - 1;
    if (C <= 8) {
      // shared memory per trajectory: two raw-rollout buffers + the staged signature;
      // ~73 KB per CTA keeps three persistent CTAs on an SM (two for C >= 7, whose
      // C*C accumulators need more than 80 registers); the grid is exactly the number
      // of co-resident CTAs
      const int64_t budget = (C >= 7 ? 110 : 73) * 1024;
      const int64_t per_traj = (2 * len * (d + a) + p.siglen) * 4;
      BSIG_REQUIRE(per_traj <= 200 * 1024, "signature: path too long for shared memory");
      // seven consumer warps + the producer warp = 256 threads
      int tpb = (int)std::min<int64_t>(224 / C, std::max<int64_t>(budget / per_traj, 1));
      if (const char* e = getenv("BSIG_SIG_TPB")) tpb = std::max(1, std::min(tpb, atoi(e)));
      if (tpb >= 4) tpb &= ~3;              // multiples of 4 keep every tile 16-byte aligned
      p.tpb = tpb;
      p.raw_stride = (int)((tpb * len * (d + a) + 3) & ~(int64_t)3);
      auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
      p.bulk = (tpb % 4 == 0) && t_states == len && (a == 0 || t_actions == len) &&
               al16(states) && (a == 0 || al16(actions)) && al16(out);
      const size_t smem = (size_t)(2 * p.raw_stride + tpb * p.siglen) * 4;
      const int threads = (int)(ceil_div((int64_t)tpb * C, 32) * 32) + 32;   // + producer warp
      const int64_t ntiles = ceil_div(n, tpb);
      // packed-FFMA2 rows: FFMA2 issues on the fma-heavy pipe only, FFMA on either, so
      // the split is a throughput knob (measured)
      static const int jp_env = [] {
        const char* e = getenv("BSIG_SIG_JP");
        return e ? atoi(e) : -1;
      }();
#define BSIG_SIG_LAUNCH(CV, JPV)                                                             \
  {                                                                                          \
    auto kern = signature3_small_kernel<CV, JPV>;                                            \
    if (smem > 48 * 1024)                                                                    \
      BSIG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,      \
                                     (int)smem));                                            \
    int occ = 1;                                                                             \
    BSIG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem));     \
    const unsigned grid =                                                                    \
        (unsigned)std::min<int64_t>(ntiles, std::max(occ, 1) * (int64_t)sm_count());         \
    kern<<<grid, threads, smem, st>>>(p);                                                    \
  }
#define BSIG_SIGS(CV)                                                                        \
  case CV: {                                                                                 \
    const int jp = jp_env < 0 ? CV / 2 : jp_env;                                             \
    if (jp <= 0) BSIG_SIG_LAUNCH(CV, 0)                                                      \
    else if (jp == 1) BSIG_SIG_LAUNCH(CV, 1)                                                 \
    else if (jp < CV) BSIG_SIG_LAUNCH(CV, CV / 2)                                            \
    else BSIG_SIG_LAUNCH(CV, CV)                                                             \
  } break;
      switch ((int)C) {
        BSIG_SIGS(2) BSIG_SIGS(3) BSIG_SIGS(4) BSIG_SIGS(5) BSIG_SIGS(6) BSIG_SIGS(7) BSIG_SIGS(8)
      }
#undef BSIG_SIG_LAUNCH
#undef BSIG_SIGS
    } else {
      const int cc = (int)(C * C);
      const int cpad = C <= 16 ? 16 : 24;
      int tpb = std::max(1, 256 / cc);
      const int64_t per_traj = (steps * cpad + std::max<int64_t>(p.siglen, len * (d + a))) * 4;
      tpb = (int)std::max<int64_t>(1, std::min<int64_t>(tpb, (100 * 1024) / per_traj));
      p.tpb = tpb;
      const int threads = (int)(ceil_div((int64_t)tpb * cc, 32) * 32);
      const size_t smem = (size_t)tpb * per_traj;
      BSIG_REQUIRE(smem <= 200 * 1024, "signature: path too long for shared memory");
      const unsigned grid = (unsigned)ceil_div(n, tpb);
      if (cpad == 16) {
        if (smem > 48 * 1024)
          BSIG_CUDA(cudaFuncSetAttribute(signature3_kernel<16>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        signature3_kernel<16><<<grid, threads, smem, st>>>(p);
      } else {
        if (smem > 48 * 1024)
          BSIG_CUDA(cudaFuncSetAttribute(signature3_kernel<24>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        signature3_kernel<24><<<grid, threads, smem, st>>>(p);
      }
    }
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

// ===================================================================== backward
// Reverse sweep of the Chen recursion.  With S' = S (x) exp(d) and adjoints
// (G1', G2', G3) of the levels after a step (G3 never changes), the adjoint of
// the increment is
//   dd[a] = G1'[a] + 1/2 sum_j G2'[a,j] d_j + 1/6 sum_jk G3[a,j,k] d_j d_k          (own row)
//         + sum_i { G2'[i,a] (S1_i + d_i/2) + sum_j G3[i,j,a] (S2_ij + S1_i d_j/2 + d_i d_j/6)
//                   + (S1_i/2 + d_i/6) sum_k G3[i,a,k] d_k }                        (all rows)
// and the adjoints before the step are
//   G2[i,j] = G2'[i,j] + sum_k G3[i,j,k] d_k,
//   G1[i]   = G1'[i] + sum_j G2'[i,j] d_j + 1/2 sum_jk G3[i,j,k] d_j d_k.
// S1 before a step is the prefix sum of increments; S2 before a step is recovered
// by undoing the forward update (S2 -= (S1 + d/2) (x) d), so nothing is stored.
// The path gradient is dx_{t+1} += dd_t, dx_t -= dd_t (time channel dropped).
namespace bsig {

struct SigBwdArgs {
  const float* states;
  const float* actions;
  const float* grad;        // [n, siglen]
  float* d_states;          // [n, L, D]
  float* d_actions;         // [n, L, A]
  int64_t n;
  int64_t s_stride, a_stride;
  int L, D, A, C;
  int tpb;
  int64_t siglen;
};

__device__ __forceinline__ void store_path_grad(const SigBwdArgs& p, int64_t traj, int t, int c,
                                                float v) {
  if (c == 0) return;                                   // time channel
  if (c <= p.D) p.d_states[(traj * p.L + t) * p.D + (c - 1)] = v;
  else p.d_actions[(traj * p.L + t) * p.A + (c - 1 - p.D)] = v;
}

// depth 3, C <= 8: thread i of a trajectory owns row i of every level.  The C
// threads of a trajectory are consecutive lanes of one warp; the per-step sum
// over rows goes through a warp-private shared-memory tile.
template <int C>
__global__ void __launch_bounds__(256) signature3_bwd_small_kernel(SigBwdArgs p) {
  extern __shared__ __align__(16) float smem[];
  constexpr int CPAD = 8;
  constexpr int TPW = 32 / C;                 // trajectories per warp
  const int steps = p.L - 1;
  const int tstride = steps * CPAD + 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  float* ds = smem;                                        // [tpb][tstride]
  float* xch = smem + (size_t)p.tpb * tstride;             // [nwarp][TPW][C][C] exchange tiles
  float* raw = xch + (size_t)nwarp * TPW * C * C;          // raw rollouts (load_increments)
  const int64_t traj0 = (int64_t)blockIdx.x * p.tpb;
  const int ntraj = (int)min((int64_t)p.tpb, p.n - traj0);
  SigArgs q;
  q.states = p.states; q.actions = p.actions; q.out = nullptr; q.n = p.n;
  q.s_stride = p.s_stride; q.a_stride = p.a_stride; q.L = p.L; q.D = p.D; q.A = p.A; q.C = p.C;
  q.tpb = p.tpb; q.siglen = p.siglen;
  load_increments<CPAD>(q, ds, tstride, raw, traj0, ntraj);
  __syncthreads();

  const int tw = lane / C, i = lane - tw * C;              // trajectory within the warp, row
  const int tl = warp * TPW + tw;
  const bool active = (tw < TPW) && (tl < ntraj);
  const int tls = active ? tl : 0;
  const float* d = ds + tls * tstride;
  const int64_t traj = traj0 + tls;
  const float* gr = p.grad + traj * p.siglen;
  float* x = xch + ((size_t)warp * TPW + (tw < TPW ? tw : 0)) * C * C;

  float g1 = active ? __ldg(gr + i) : 0.f;
  float g2[C], g3[C][C], s2[C];
  float s1 = 0.f;
#pragma unroll
  for (int j = 0; j < C; ++j) {
    g2[j] = active ? __ldg(gr + C + i * C + j) : 0.f;
    s2[j] = 0.f;
#pragma unroll
    for (int k = 0; k < C; ++k) g3[j][k] = active ? __ldg(gr + C + C * C + (i * C + j) * C + k) : 0.f;
  }
  // forward pass for S2 (row i) and the total S1
  for (int t = 0; t < steps; ++t) {
    const float di = d[t * CPAD + i];
    const float a2 = s1 + di * 0.5f;
#pragma unroll
    for (int j = 0; j < C; ++j) s2[j] = fmaf(a2, d[t * CPAD + j], s2[j]);
    s1 += di;
  }
  float dd_next = 0.f;                                     // dd of step t+1 (for dx_{t+1})
  for (int t = steps - 1; t >= 0; --t) {
    const float4 lo = *reinterpret_cast<const float4*>(d + t * CPAD);
    const float4 hi = *reinterpret_cast<const float4*>(d + t * CPAD + 4);
    const float dv[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
    const float di = d[t * CPAD + i];
    s1 -= di;                                              // S1 before this step
    const float a2 = s1 + di * 0.5f;
    const float u = s1 * 0.5f + di * (1.0f / 6.0f);
    float own = g1, g1_add = 0.f, bsum = 0.f;
    float contrib[C];
#pragma unroll
    for (int a = 0; a < C; ++a) contrib[a] = g2[a] * a2;
#pragma unroll
    for (int j = 0; j < C; ++j) {
      s2[j] = fmaf(-a2, dv[j], s2[j]);                     // undo the forward update
      const float tij = s2[j] + (s1 * 0.5f + di * (1.0f / 6.0f)) * dv[j];
      float aij = 0.f;
#pragma unroll
      for (int k = 0; k < C; ++k) {
        aij = fmaf(g3[j][k], dv[k], aij);
        contrib[k] = fmaf(g3[j][k], tij, contrib[k]);      // sum_j G3[i,j,a] T_ij  (a = k)
      }
      contrib[j] = fmaf(u, aij, contrib[j]);               // (S1_i/2 + d_i/6) A_ia  (a = j)
      own = fmaf(0.5f * g2[j], dv[j], own);
      g1_add = fmaf(g2[j], dv[j], g1_add);
      bsum = fmaf(aij, dv[j], bsum);
      g2[j] += aij;                                        // adjoint of S2 before the step
    }
    own = fmaf(bsum, 1.0f / 6.0f, own);
    g1 += g1_add + 0.5f * bsum;
    // dd[a] = own(a) + sum_i contrib_i[a]
    __syncwarp();
    if (tw < TPW) {
#pragma unroll
      for (int a = 0; a < C; ++a) x[i * C + a] = contrib[a];
    }
    __syncwarp();
    float dd = own;
#pragma unroll
    for (int r = 0; r < C; ++r) dd += x[r * C + i];
    if (active) store_path_grad(p, traj, t + 1, i, dd - dd_next);
    dd_next = dd;
  }
  if (active) store_path_grad(p, traj, 0, i, -dd_next);
}

// depth 3, 9 <= C <= 22 (the widest path that still gets depth 3): one CTA per trajectory,
// thread (i, j) owns entry (i, j) of the level-2 adjoint G2 and of S2; the C^3 level-3
// adjoint G3 (constant through the sweep) sits in shared memory.  Per reverse step
//   A_ij = sum_k G3[i,j,k] d_k                      (thread (i,j), C FMAs)
//   T_ij = S2_ij + (S1_i/2 + d_i/6) d_j             (before the step)
//   dd[a] = G1'[a] + sum_j (G2'[a,j]/2 + A_aj/6) d_j                                (row a)
//         + sum_i { G2'[i,a] (S1_i + d_i/2) + (S1_i/2 + d_i/6) A_ia }               (column a)
//         + sum_ij G3[i,j,a] T_ij                  (thread (i,a): sum over j, then over i)
//   G2 = G2' + A,   G1[i] = G1'[i] + sum_j (G2'[i,j] + A_ij/2) d_j.
// Same formulas as signature3_bwd_small_kernel, with the row / column sums going through
// shared memory instead of registers.
__global__ void __launch_bounds__(512) signature3_bwd_kernel(SigBwdArgs p) {
  extern __shared__ __align__(16) float smem[];
  const int C = p.C, L = p.L, steps = L - 1, CC = C * C;
  float* x = smem;                              // [L][C] path points
  float* g3 = x + (size_t)L * C;                // [C][C][C]
  float* tm = g3 + (size_t)CC * C;              // [C][C]  T_ij
  float* rw = tm + CC;                          // [C][C]  row terms of dd
  float* rg = rw + CC;                          // [C][C]  row terms of the G1 update
  float* cl = rg + CC;                          // [C][C]  column terms of dd
  float* p3 = cl + CC;                          // [C][C]  p3[i][a] = sum_j G3[i,j,a] T_ij
  float* g1s = p3 + CC;                         // [C]
  float* dv = g1s + C;                          // [C] increment of the current step
  const int64_t traj = blockIdx.x;
  const int tid = threadIdx.x;
  SigArgs q;
  q.states = p.states; q.actions = p.actions; q.out = nullptr; q.n = p.n;
  q.s_stride = p.s_stride; q.a_stride = p.a_stride; q.L = L; q.D = p.D; q.A = p.A; q.C = C;
  q.tpb = 1; q.siglen = p.siglen;
  load_paths(q, x, traj, 1);
  const float* gr = p.grad + traj * p.siglen;
  for (int e = tid; e < CC * C; e += blockDim.x) g3[e] = __ldg(gr + C + CC + e);
  for (int e = tid; e < C; e += blockDim.x) g1s[e] = __ldg(gr + e);
  const bool live = tid < CC;
  const int i = live ? tid / C : 0, j = live ? tid - i * C : 0;
  float g2 = live ? __ldg(gr + C + tid) : 0.f;
  __syncthreads();
  // forward sweep: S2_ij and S1_i after the last step
  float s1 = 0.f, s2 = 0.f;
  for (int t = 0; t < steps; ++t) {
    const float di = x[(t + 1) * C + i] - x[t * C + i];
    const float dj = x[(t + 1) * C + j] - x[t * C + j];
    s2 = fmaf(s1 + 0.5f * di, dj, s2);
    s1 += di;
  }
  float dd_next = 0.f;
  for (int t = steps - 1; t >= 0; --t) {
    if (tid < C) dv[tid] = x[(t + 1) * C + tid] - x[t * C + tid];
    __syncthreads();
    if (live) {
      const float di = dv[i], dj = dv[j];
      s1 -= di;                                           // S1_i before this step
      const float a2 = s1 + 0.5f * di;
      const float u = 0.5f * s1 + di * (1.0f / 6.0f);
      s2 = fmaf(-a2, dj, s2);                             // S2_ij before this step
      const float* row = g3 + (size_t)tid * C;
      float aij = 0.f;
      for (int k = 0; k < C; ++k) aij = fmaf(row[k], dv[k], aij);
      tm[tid] = fmaf(u, dj, s2);
      rw[tid] = (0.5f * g2 + aij * (1.0f / 6.0f)) * dj;
      rg[tid] = (g2 + 0.5f * aij) * dj;
      cl[tid] = fmaf(g2, a2, u * aij);
      g2 += aij;
    }
    __syncthreads();
    if (live) {                                           // thread (i, a = j)
      float acc = 0.f;
      const float* g3i = g3 + (size_t)i * CC + j;
      const float* ti = tm + i * C;
      for (int jj = 0; jj < C; ++jj) acc = fmaf(g3i[jj * C], ti[jj], acc);
      p3[tid] = acc;
    }
    __syncthreads();
    if (tid < C) {
      const int a = tid;
      float own = g1s[a], upd = 0.f, col = 0.f;
      for (int q2 = 0; q2 < C; ++q2) {
        own += rw[a * C + q2];
        upd += rg[a * C + q2];
        col += cl[q2 * C + a] + p3[q2 * C + a];
      }
      const float dd = own + col;
      g1s[a] += upd;
      store_path_grad(p, traj, t + 1, a, dd - dd_next);
      dd_next = dd;
    }
    // the next step's dv write happens after its own barrier; tm / rw / rg / cl / p3 are
    // rewritten only after that barrier as well
  }
  if (tid < C) store_path_grad(p, traj, 0, tid, -dd_next);
}

// depth 1 and 2, any C: one CTA per trajectory.  At depth 2 the level-2 adjoint is
// constant (G2) and G1 before step t is G1 + G2 (x_{L-1} - x_{t+1}).
__global__ void __launch_bounds__(256) signature12_bwd_kernel(SigBwdArgs p, int depth) {
  extern __shared__ float smem[];
  const int C = p.C, L = p.L;
  float* x = smem;                              // [L][C]
  float* g2 = smem + (size_t)L * C;             // [C][C] (depth 2)
  const int64_t traj = blockIdx.x;
  SigArgs q;
  q.states = p.states; q.actions = p.actions; q.out = nullptr; q.n = p.n;
  q.s_stride = p.s_stride; q.a_stride = p.a_stride; q.L = L; q.D = p.D; q.A = p.A; q.C = C;
  q.tpb = 1; q.siglen = p.siglen;
  load_paths(q, x, traj, 1);
  const float* gr = p.grad + traj * p.siglen;
  if (depth >= 2)
    for (int e = threadIdx.x; e < C * C; e += blockDim.x) g2[e] = __ldg(gr + C + e);
  __syncthreads();
  for (int a = threadIdx.x; a < C; a += blockDim.x) {
    const float g1 = __ldg(gr + a);
    float dd_next = 0.f;
    for (int t = L - 2; t >= 0; --t) {
      float dd = g1;
      if (depth >= 2) {
        const float* x0 = x + t * C;
        const float* x1 = x0 + C;
        const float* xl = x + (L - 1) * C;
        float acc = 0.f;
        for (int j = 0; j < C; ++j) {
          const float dj = x1[j] - x0[j];
          // row a:  G2[a,j] ((x_last - x_{t+1})_j + d_j/2);  column a: G2[j,a] (S1_j + d_j/2)
          acc = fmaf(g2[a * C + j], (xl[j] - x1[j]) + 0.5f * dj, acc);
          acc = fmaf(g2[j * C + a], (x0[j] - x[j]) + 0.5f * dj, acc);
        }
        dd += acc;
      }
      store_path_grad(p, traj, t + 1, a, dd - dd_next);
      dd_next = dd;
    }
    store_path_grad(p, traj, 0, a, -dd_next);
  }
}

}  // namespace bsig

extern "C" int bsig_signature_bwd(const float* states, const float* actions, const float* grad_out,
                                  float* d_states, float* d_actions, int64_t n, int64_t len,
                                  int64_t t_states, int64_t t_actions, int64_t d, int64_t a,
                                  int depth, void* stream) {
  BSIG_REQUIRE(n >= 0 && d >= 1 && a >= 0, "signature_bwd: bad sizes");
  BSIG_REQUIRE(len >= 2 && t_states >= len && (a == 0 || t_actions >= len),
               "signature_bwd: need >= 2 path points and len <= stored steps");
  BSIG_REQUIRE(depth >= 1 && depth <= 3, "signature_bwd: depth must be 1, 2 or 3");
  if (n == 0) return 0;
  SigBwdArgs p;
  p.states = states; p.actions = actions; p.grad = grad_out;
  p.d_states = d_states; p.d_actions = d_actions; p.n = n;
  p.s_stride = t_states * d; p.a_stride = t_actions * a;
  p.L = (int)len; p.D = (int)d; p.A = (int)a; p.C = (int)(1 + d + a);
  const int64_t C = p.C;
  p.siglen = C + (depth >= 2 ? C * C : 0) + (depth >= 3 ? C * C * C : 0);
  cudaStream_t st = (cudaStream_t)stream;
  if (depth == 3) {
    BSIG_REQUIRE(C <= 22, "signature_bwd: depth 3 supports at most 22 channels (got %d)", (int)C);
    if (C > 8) {
      // wide paths: one CTA per trajectory, G3 in shared memory
      p.tpb = 1;
      const size_t smem = ((size_t)len * C + (size_t)C * C * C + 5 * (size_t)C * C + 2 * (size_t)C) * 4;
      BSIG_REQUIRE(smem <= 200 * 1024, "signature_bwd: path too long for shared memory");
      if (smem > 48 * 1024)
        BSIG_CUDA(cudaFuncSetAttribute(signature3_bwd_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      const int threads = (int)((C * C + 31) / 32 * 32);
      signature3_bwd_kernel<<<(unsigned)n, threads, smem, st>>>(p);
      BSIG_LAUNCH_CHECK();
      return 0;
    }
    const int tpw = 32 / (int)C, nwarp = 8;
    p.tpb = tpw * nwarp;
    const int64_t steps = len - 1;
    const size_t smem = ((size_t)p.tpb * (steps * 8 + 8) + (size_t)nwarp * tpw * C * C +
                         (size_t)p.tpb * len * (d + a)) * 4;
    BSIG_REQUIRE(smem <= 200 * 1024, "signature_bwd: path too long for shared memory");
    const unsigned grid = (unsigned)ceil_div(n, p.tpb);
#define BSIG_SIGB(CV)                                                                        \
  case CV:                                                                                   \
    if (smem > 48 * 1024)                                                                    \
      BSIG_CUDA(cudaFuncSetAttribute(signature3_bwd_small_kernel<CV>,                        \
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    signature3_bwd_small_kernel<CV><<<grid, 256, smem, st>>>(p);                             \
    break;
    switch ((int)C) {
      BSIG_SIGB(2) BSIG_SIGB(3) BSIG_SIGB(4) BSIG_SIGB(5) BSIG_SIGB(6) BSIG_SIGB(7) BSIG_SIGB(8)
    }
#undef BSIG_SIGB
  } else {
    p.tpb = 1;
    const size_t smem = ((size_t)len * C + (depth >= 2 ? (size_t)C * C : 0)) * 4;
    BSIG_REQUIRE(smem <= 200 * 1024, "signature_bwd: path too large for shared memory");
    if (smem > 48 * 1024)
      BSIG_CUDA(cudaFuncSetAttribute(signature12_bwd_kernel,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    signature12_bwd_kernel<<<(unsigned)n, 256, smem, st>>>(p, depth);
  }
  BSIG_LAUNCH_CHECK();
  return 0;
}
