// Shared helpers for the bsig kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "bsig.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "bsig kernels are written for sm_100a (B200) only"
#endif

namespace bsig {

void set_error(const char* fmt, ...);
int sm_count();
void count_launch();
bool pdl_enabled();

// Programmatic dependent launch (PDL): the step kernels of a training call form a
// chain of short, dependent launches.  Each of them (1) does its operand-independent
// set-up, (2) waits for the previous kernel in the stream to complete and flush
// (griddepcontrol.wait), (3) immediately allows the NEXT kernel to be scheduled
// (griddepcontrol.launch_dependents).  Because (3) comes after (2), kernel N+1 can only
// become resident once kernel N-1 has finished, so N+1's set-up overlaps N's execution
// and nothing N-1 or earlier wrote is read early.  Both instructions are no-ops when the
// launch did not carry the programmatic-serialization attribute.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait_then_release() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
#endif
// Appends the PDL launch attribute (if enabled) to attrs[n]; returns the new count.
static inline int add_pdl_attr(cudaLaunchAttribute* attrs, int n) {
  if (!pdl_enabled()) return n;
  attrs[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attrs[n].val.programmaticStreamSerializationAllowed = 1;
  return n + 1;
}

#define BSIG_REQUIRE(cond, ...)            \
  do {                                     \
    if (!(cond)) {                         \
      ::bsig::set_error(__VA_ARGS__);      \
      return 1;                            \
    }                                      \
  } while (0)

#define BSIG_CUDA(call)                                                        \
  do {                                                                         \
    cudaError_t e__ = (call);                                                  \
    if (e__ != cudaSuccess) {                                                  \
      ::bsig::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call,           \
                        cudaGetErrorString(e__));                              \
      return 2;                                                                \
    }                                                                          \
  } while (0)

#define BSIG_LAUNCH_CHECK()            \
  do {                                 \
    ::bsig::count_launch();            \
    BSIG_CUDA(cudaGetLastError());     \
  } while (0)

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Deterministic block-wide sum; result valid in every thread.  `scratch` holds
// at least 33 elements.  All threads of the block must call it.
template <typename T>
__device__ __forceinline__ T block_sum(T v, T* scratch) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) scratch[wid] = v;
  __syncthreads();
  if (wid == 0) {
    T t = lane < nw ? scratch[lane] : T(0);
    t = warp_sum(t);
    if (lane == 0) scratch[32] = t;
  }
  __syncthreads();
  return scratch[32];
}

// Two sums at once (same fixed order per component as block_sum).
__device__ __forceinline__ float2 block_sum2(float2 v, float2* scratch) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    v.x += __shfl_xor_sync(0xffffffffu, v.x, o);
    v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
  }
  __syncthreads();
  if (lane == 0) scratch[wid] = v;
  __syncthreads();
  if (wid == 0) {
    float2 t = lane < nw ? scratch[lane] : make_float2(0.f, 0.f);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      t.x += __shfl_xor_sync(0xffffffffu, t.x, o);
      t.y += __shfl_xor_sync(0xffffffffu, t.y, o);
    }
    if (lane == 0) scratch[32] = t;
  }
  __syncthreads();
  return scratch[32];
}

// Streaming (evict-first) 128-bit store: summaries are written once and read
// much later, keep them out of the way of the L2-resident inputs.
__device__ __forceinline__ void st_stream_f4(float* p, float4 v) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w)
               : "memory");
}

__device__ __forceinline__ bool finite_f(float v) { return fabsf(v) <= 3.402823466e38f; }

}  // namespace bsig
