// Profiling aid (not on the product path): a persistent copy kernel built from exactly the
// bulk-copy pipeline of the streaming kernels (cp.async.bulk global->shared on an mbarrier
// double buffer, cp.async.bulk shared->global as bulk groups, one producer thread), with
// no arithmetic.  It answers "what does the 1-D bulk path sustain on this machine?", the
// ceiling signature3_small_kernel / nll_stream_kernel are measured against
// (profiles/bulk_probe.py).
#include <algorithm>

#include "async_copy.cuh"
#include "common.cuh"

namespace bsig {

__global__ void __launch_bounds__(128)
bulk_copy_probe_kernel(const char* __restrict__ src, char* __restrict__ dst, int64_t ntiles,
                       int tile_bytes, int stages) {
  extern __shared__ __align__(16) char smem[];
  __shared__ __align__(8) uint64_t full[8];
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int s = 0; s < stages; ++s) ac::mbar_init(&full[s], 1);
    ac::fence_barrier_init();
  }
  __syncthreads();
  if (tid != 0) return;
  const int my_tiles = (int64_t)blockIdx.x < ntiles
                           ? (int)((ntiles - 1 - blockIdx.x) / gridDim.x) + 1 : 0;
  auto tile_of = [&](int it) { return (int64_t)blockIdx.x + (int64_t)it * gridDim.x; };
  auto load = [&](int it) {
    const int s = it % stages;
    ac::mbar_expect_tx(&full[s], (uint32_t)tile_bytes);
    ac::bulk_g2s(smem + (size_t)s * tile_bytes, src + tile_of(it) * tile_bytes,
                 (uint32_t)tile_bytes, &full[s]);
  };
  for (int it = 0; it < min(my_tiles, stages - 1); ++it) load(it);
  for (int it = 0; it < my_tiles; ++it) {
    const int s = it % stages;
    ac::mbar_wait(&full[s], (uint32_t)(it / stages) & 1u);
    ac::bulk_s2g(dst + tile_of(it) * tile_bytes, smem + (size_t)s * tile_bytes,
                 (uint32_t)tile_bytes);
    ac::bulk_commit();
    const int nxt = it + stages - 1;            // refills the stage whose store was issued
    if (nxt < my_tiles) {                       // one iteration ago
      ac::bulk_wait_read<1>();
      load(nxt);
    }
  }
  ac::bulk_wait_read<0>();
}

}  // namespace bsig

using namespace bsig;

extern "C" int bsig_bulk_copy_probe(const void* src, void* dst, int64_t bytes, int64_t tile_bytes,
                                    int stages, int ctas_per_sm, void* stream) {
  BSIG_REQUIRE(tile_bytes >= 16 && tile_bytes % 16 == 0 && bytes % tile_bytes == 0,
               "bulk_copy_probe: sizes must be multiples of 16 bytes / of the tile");
  BSIG_REQUIRE(stages >= 2 && stages <= 8 && ctas_per_sm >= 1, "bulk_copy_probe: bad pipeline");
  const size_t smem = (size_t)stages * tile_bytes;
  BSIG_REQUIRE(smem <= 200 * 1024, "bulk_copy_probe: stages do not fit shared memory");
  if (smem > 48 * 1024)
    BSIG_CUDA(cudaFuncSetAttribute(bulk_copy_probe_kernel,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t ntiles = bytes / tile_bytes;
  const int grid = (int)std::min<int64_t>(ntiles, (int64_t)ctas_per_sm * sm_count());
  bulk_copy_probe_kernel<<<grid, 128, smem, (cudaStream_t)stream>>>(
      (const char*)src, (char*)dst, ntiles, (int)tile_bytes, stages);
  BSIG_LAUNCH_CHECK();
  return 0;
}
