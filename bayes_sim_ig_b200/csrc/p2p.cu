// Data-parallel exchange step fused with the optimiser: a one-shot all-reduce of the
// flat gradient buffer over NVLink peer memory + Adam, in ONE kernel per update.
//
// The reference has no multi-GPU path; the exchange a data-parallel
// MDNN / MDRFF needs is one sum of a 0.1-0.4 MB gradient vector per update -- pure
// latency.  Instead of an NCCL launch between two graphs, every rank's Adam kernel
//   1. publishes "my gradients of this update are complete" by storing the update's
//      epoch into every peer's flag array (st.release.sys over NVLink),
//   2. waits until all peers' epochs have arrived in its own flag array,
//   3. reads the gradient of each element from all peers' buffers (peer loads over
//      NVSwitch, summed in rank order => bit-identical on every rank) and applies Adam.
// Gradient buffers are double-buffered by update parity, so a rank can start the next
// backward pass while peers still read the previous buffer.  The epoch lives in device
// memory and is advanced by the kernel itself, so the launch is CUDA-graph capturable
// and replayable.  Buffers are cudaMalloc'ed here (IPC handles need whole allocations).
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace bsig {

constexpr int kMaxPeers = 8;

struct P2PArgs {
  const float* grads[kMaxPeers];     // gradient buffer of every rank (this update's parity)
  unsigned int* flags[kMaxPeers];    // flag array of every rank: flags[r][q] = epoch published by q
  unsigned int* ctrl;                // local: [0] epoch counter, [1] finished-block counter
  int rank, world;
};

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Watchdog: a rank that died (exception, OOM) between two updates never publishes its epoch.
// The wait is therefore bounded by %globaltimer; on expiry the waiting thread records
// 1 + (index of the missing peer) in the STICKY error word ctrl[2], every later launch skips
// its wait (fast fail instead of one time-out per update) and the host turns the flag into an
// exception after the call (train_engine.py).  Results of a timed-out call are discarded.
__global__ void __launch_bounds__(256)
adam_allreduce_kernel(P2PArgs a, float* __restrict__ p, float* __restrict__ m,
                      float* __restrict__ v, int64_t count, float one_minus_b1, float b2,
                      float one_minus_b2, float step_size, float inv_bc2_sqrt, float eps,
                      float gscale, unsigned long long timeout_ns) {
  __shared__ unsigned int s_epoch;
  pdl_wait_then_release();       // the backward kernels of this update are complete
  // ctrl[8..15]: nanosecond stamps of block 0 (start, flags published, peers arrived,
  // reduction + Adam done) of the most recent launch -- read by profiles/dp_smoke.py
  unsigned long long* stamps = reinterpret_cast<unsigned long long*>(a.ctrl + 8);
  const bool tracer = blockIdx.x == 0 && threadIdx.x == 0;
  if (tracer) stamps[0] = globaltimer_ns();
  if (threadIdx.x == 0) s_epoch = a.ctrl[0] + 1u;
  __syncthreads();
  const unsigned int epoch = s_epoch;
  if (blockIdx.x == 0 && threadIdx.x < a.world) {
    // my backward kernels precede this kernel in stream order, so their writes are
    // complete; one system-scope fence per publishing thread (fence + relaxed store = a
    // release), and the `world` flag stores travel over NVLink in parallel instead of
    // waiting for each other's acknowledgement
    __threadfence_system();
    st_relaxed_sys(a.flags[threadIdx.x] + a.rank, epoch);
  }
  if (tracer) stamps[1] = globaltimer_ns();
  if (threadIdx.x < a.world) {
    const unsigned int* mine = a.flags[a.rank] + threadIdx.x;
    volatile unsigned int* err = a.ctrl + 2;
    if (*err == 0u) {
      const unsigned long long t_begin = globaltimer_ns();
      unsigned int spins = 0;
      while ((int)(ld_acquire_sys(mine) - epoch) < 0) {   // peers are at most one update behind
        if ((++spins & 255u) == 0u && globaltimer_ns() - t_begin > timeout_ns) {
          atomicCAS(a.ctrl + 2, 0u, 1u + threadIdx.x);
          break;
        }
      }
    }
  }
  __syncthreads();
  if (tracer) stamps[2] = globaltimer_ns();

  const int64_t n4 = count / 4;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  auto upd = [&](float& pp, float gg, float& mm, float& vv) {
    gg *= gscale;
    mm = mm + (gg - mm) * one_minus_b1;
    vv = vv * b2 + one_minus_b2 * gg * gg;
    const float denom = sqrtf(vv) * inv_bc2_sqrt + eps;
    pp = pp - step_size * (mm / denom);
  };
  float4* p4 = reinterpret_cast<float4*>(p);
  float4* m4 = reinterpret_cast<float4*>(m);
  float4* v4 = reinterpret_cast<float4*>(v);
  for (int64_t i = t0; i < n4; i += stride) {
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int q = 0; q < kMaxPeers; ++q) {
      if (q < a.world) {
        const float4 t = __ldcv(reinterpret_cast<const float4*>(a.grads[q]) + i);
        g.x += t.x; g.y += t.y; g.z += t.z; g.w += t.w;
      }
    }
    float4 pp = p4[i], mm = m4[i], vv = v4[i];
    upd(pp.x, g.x, mm.x, vv.x);
    upd(pp.y, g.y, mm.y, vv.y);
    upd(pp.z, g.z, mm.z, vv.z);
    upd(pp.w, g.w, mm.w, vv.w);
    p4[i] = pp; m4[i] = mm; v4[i] = vv;
  }
  for (int64_t i = n4 * 4 + t0; i < count; i += stride) {
    float g = 0.f;
    for (int q = 0; q < a.world; ++q) g += __ldcv(a.grads[q] + i);
    upd(p[i], g, m[i], v[i]);
  }

  // the last block to finish advances the epoch for the next update
  __syncthreads();
  if (tracer) stamps[3] = globaltimer_ns();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int done = atomicAdd(a.ctrl + 1, 1u);
    if (done == gridDim.x - 1) {
      a.ctrl[1] = 0u;
      __threadfence();
      a.ctrl[0] = epoch;
    }
  }
}

}  // namespace bsig

using namespace bsig;

static unsigned long long g_p2p_timeout_ns = [] {
  const char* e = getenv("BSIG_P2P_TIMEOUT_MS");
  const long long ms = e ? atoll(e) : 20000;
  return (unsigned long long)(ms < 1 ? 1 : ms) * 1000000ull;
}();

extern "C" int bsig_p2p_set_timeout_ms(int64_t ms) {
  BSIG_REQUIRE(ms >= 1, "p2p_set_timeout_ms: time-out must be positive");
  g_p2p_timeout_ns = (unsigned long long)ms * 1000000ull;
  return 0;
}

extern "C" int bsig_p2p_alloc(void** ptr, int64_t bytes, unsigned char* handle64) {
  BSIG_REQUIRE(ptr != nullptr && handle64 != nullptr && bytes > 0, "p2p_alloc: bad arguments");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  void* p = nullptr;
  BSIG_CUDA(cudaMalloc(&p, (size_t)bytes));
  BSIG_CUDA(cudaMemset(p, 0, (size_t)bytes));
  BSIG_CUDA(cudaDeviceSynchronize());
  cudaIpcMemHandle_t h;
  BSIG_CUDA(cudaIpcGetMemHandle(&h, p));
  memcpy(handle64, &h, 64);
  *ptr = p;
  return 0;
}

extern "C" int bsig_p2p_open(const unsigned char* handle64, void** ptr) {
  BSIG_REQUIRE(ptr != nullptr && handle64 != nullptr, "p2p_open: bad arguments");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  BSIG_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}

extern "C" int bsig_p2p_read(const void* dev_ptr, void* host_ptr, int64_t bytes) {
  BSIG_CUDA(cudaMemcpy(host_ptr, dev_ptr, (size_t)bytes, cudaMemcpyDeviceToHost));
  return 0;
}

extern "C" int bsig_p2p_close(void* ptr) {
  BSIG_CUDA(cudaIpcCloseMemHandle(ptr));
  return 0;
}

extern "C" int bsig_p2p_free(void* ptr) {
  BSIG_CUDA(cudaFree(ptr));
  return 0;
}

extern "C" int bsig_adam_allreduce_step(float* param, const void* const* peer_grads,
                                        void* const* peer_flags, void* ctrl, int rank, int world,
                                        float* exp_avg, float* exp_avg_sq, int64_t count,
                                        int64_t step, float lr, float beta1, float beta2, float eps,
                                        void* stream) {
  BSIG_REQUIRE(world >= 1 && world <= kMaxPeers && rank >= 0 && rank < world,
               "adam_allreduce: world must be 1..%d", kMaxPeers);
  BSIG_REQUIRE(count >= 1 && step >= 1, "adam_allreduce: bad count/step");
  P2PArgs a;
  for (int q = 0; q < kMaxPeers; ++q) {
    a.grads[q] = q < world ? (const float*)peer_grads[q] : nullptr;
    a.flags[q] = q < world ? (unsigned int*)peer_flags[q] : nullptr;
  }
  a.ctrl = (unsigned int*)ctrl;
  a.rank = rank;
  a.world = world;
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  const float step_size = (float)((double)lr / bc1);
  const float inv_bc2_sqrt = (float)(1.0 / sqrt(bc2));
  // every block must be resident at once (blocks spin on the peers' flags)
  const int blocks =
      (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(count / 4 + 1, 256), sm_count()));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)blocks);
  cfg.blockDim = dim3(256);
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  cfg.attrs = attr;
  cfg.numAttrs = add_pdl_attr(attr, 0);
  BSIG_CUDA(cudaLaunchKernelEx(&cfg, adam_allreduce_kernel, a, param, exp_avg, exp_avg_sq, count,
                               1.0f - beta1, beta2, 1.0f - beta2, step_size, inv_bc2_sqrt, eps,
                               1.0f / (float)world, g_p2p_timeout_ns));
  BSIG_LAUNCH_CHECK();
  return 0;
}

// Loads adam_allreduce_kernel's code without launching it.  A kernel's FIRST launch loads its
// module lazily, which is not allowed while a stream is capturing; the training engine's warm-up
// therefore runs update 0 once eagerly -- for the exchange kernel that would mean a real
// rendezvous with the peers (and an epoch that must stay in step on every rank), so it is
// preloaded instead and the warm-up stays rank-local.
extern "C" int bsig_p2p_preload(void) {
  cudaFuncAttributes attr;
  BSIG_CUDA(cudaFuncGetAttributes(&attr, bsig::adam_allreduce_kernel));
  return 0;
}
