// What a dependent hop costs on B200 (round 2): inputs for a whole-GPU dataflow
// version of the minibatch training step.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hop hop.cu && ./hop
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdio.h>
namespace cg = cooperative_groups;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("ERR %s line %d: %s\n", #x, __LINE__, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ void st_rel(unsigned* p, unsigned v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned ld_acq(const unsigned* p) { unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }

// two CTAs (different SMs) ping-pong a flag through L2: cycles per one-way hop
__global__ void k_pingpong(unsigned* flags, int iters, long long* out) {
  if (threadIdx.x != 0) return;
  const int me = blockIdx.x;
  long long t0 = clock64();
  for (int i = 1; i <= iters; ++i) {
    if (me == 0) { st_rel(flags + 0, i); while (ld_acq(flags + 32) < (unsigned)i) {} }
    else         { while (ld_acq(flags + 0) < (unsigned)i) {} st_rel(flags + 32, i); }
  }
  if (me == 0) out[0] = clock64() - t0;
}

// producer CTA writes `n` floats then a flag; consumer CTA waits for the flag, reads the data and answers
__global__ void k_data_hop(unsigned* flags, float* buf, int n, int iters, long long* out, float* sink) {
  const int me = blockIdx.x;
  float acc = 0.f;
  long long t0 = clock64();
  for (int i = 1; i <= iters; ++i) {
    if (me == 0) {
      for (int e = threadIdx.x; e < n; e += blockDim.x) buf[e] = (float)(i + e);
      __threadfence();
      __syncthreads();
      if (threadIdx.x == 0) { st_rel(flags + 0, i); while (ld_acq(flags + 32) < (unsigned)i) {} }
      __syncthreads();
    } else {
      if (threadIdx.x == 0) while (ld_acq(flags + 0) < (unsigned)i) {}
      __syncthreads();
      for (int e = threadIdx.x; e < n; e += blockDim.x) acc += __ldcg(buf + e);
      __syncthreads();
      if (threadIdx.x == 0) st_rel(flags + 32, i);
    }
  }
  if (me == 0 && threadIdx.x == 0) out[0] = clock64() - t0;
  if (acc == 123.f) sink[0] = acc;
}

// all CTAs of a co-resident grid: barrier through one atomic counter + a generation flag
__global__ void k_grid_barrier(unsigned* ctr, unsigned* gen, int iters, long long* out) {
  long long t0 = clock64();
  for (int i = 1; i <= iters; ++i) {
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      if (atomicAdd(ctr, 1u) == gridDim.x - 1) { *ctr = 0u; __threadfence(); st_rel(gen, i); }
      else while (ld_acq(gen) < (unsigned)i) {}
    }
    __syncthreads();
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = clock64() - t0;
}
__global__ void k_coop_sync(int iters, long long* out) {
  cg::grid_group grid = cg::this_grid();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) grid.sync();
  if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = clock64() - t0;
}
__global__ void k_empty(float* p) { if (p != nullptr && threadIdx.x == 1024) p[0] = 1.f; }
__global__ void k_empty_pdl(float* p) {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (p != nullptr && threadIdx.x == 1024) p[0] = 1.f;
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const double ghz = prop.clockRate * 1e-6;
  printf("device %s, %d SMs, %.3f GHz\n", prop.name, prop.multiProcessorCount, ghz);
  unsigned* flags; long long* out; float* buf; float* sink;
  CK(cudaMalloc(&flags, 4096)); CK(cudaMalloc(&out, 64)); CK(cudaMalloc(&buf, 1 << 20)); CK(cudaMalloc(&sink, 64));
  long long h;
  const int IT = 2000;
  CK(cudaMemset(flags, 0, 4096));
  k_pingpong<<<2, 32>>>(flags, IT, out);
  CK(cudaDeviceSynchronize()); CK(cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost));
  printf("flag ping-pong between two CTAs: %.0f cycles = %.2f us per one-way hop\n", h / (2.0 * IT), h / (2.0 * IT) / ghz * 1e-3);
  for (int n : {1024, 12800, 32768}) {
    CK(cudaMemset(flags, 0, 4096));
    k_data_hop<<<2, 512>>>(flags, buf, n, IT / 4, out, sink);
    CK(cudaDeviceSynchronize()); CK(cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost));
    printf("hop carrying %6d floats (write, fence, flag, read, answer): %.0f cycles = %.2f us per round trip\n", n,
           h / (IT / 4.0), h / (IT / 4.0) / ghz * 1e-3);
  }
  for (int ctas : {16, 74, 148}) {
    CK(cudaMemset(flags, 0, 4096));
    k_grid_barrier<<<ctas, 256>>>(flags, flags + 64, IT, out);
    CK(cudaDeviceSynchronize()); CK(cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost));
    printf("atomic-counter grid barrier, %3d CTAs x 256 thr: %.0f cycles = %.2f us\n", ctas, (double)h / IT, (double)h / IT / ghz * 1e-3);
    int iters = IT; void* args[] = {&iters, &out};
    CK(cudaLaunchCooperativeKernel((void*)k_coop_sync, dim3(ctas), dim3(256), args, 0, 0));
    CK(cudaDeviceSynchronize()); CK(cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost));
    printf("cooperative grid.sync(),          %3d CTAs x 256 thr: %.0f cycles = %.2f us\n", ctas, (double)h / IT, (double)h / IT / ghz * 1e-3);
  }
  // dependent (empty) kernels inside one CUDA graph: what a kernel boundary costs
  cudaStream_t st; CK(cudaStreamCreate(&st));
  for (int pdl = 0; pdl < 2; ++pdl) {
    for (int ctas : {1, 80}) {
      cudaGraph_t graph; cudaGraphExec_t exec;
      CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed));
      for (int i = 0; i < 1000; ++i) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(ctas); cfg.blockDim = dim3(256); cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr; cfg.numAttrs = pdl;
        float* nullp = nullptr;
        if (pdl) { CK(cudaLaunchKernelEx(&cfg, k_empty_pdl, nullp)); } else { CK(cudaLaunchKernelEx(&cfg, k_empty, nullp)); }
      }
      CK(cudaStreamEndCapture(st, &graph));
      CK(cudaGraphInstantiate(&exec, graph, 0));
      CK(cudaGraphLaunch(exec, st)); CK(cudaStreamSynchronize(st));
      cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
      CK(cudaEventRecord(e0, st));
      for (int r = 0; r < 5; ++r) CK(cudaGraphLaunch(exec, st));
      CK(cudaEventRecord(e1, st)); CK(cudaEventSynchronize(e1));
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      printf("chain of 1000 dependent empty kernels in a graph, %2d CTAs, PDL %s: %.2f us per kernel\n", ctas,
             pdl ? "on " : "off", ms * 1e3 / 5000.0);
      CK(cudaGraphExecDestroy(exec)); CK(cudaGraphDestroy(graph));
    }
  }
  return 0;
}
