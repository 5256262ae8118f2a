// Error reporting and device queries of the C ABI.
#include <stdarg.h>
#include <stdlib.h>

#include <atomic>

#include "common.cuh"

namespace bsig {

static thread_local char g_err[1024] = "";
static std::atomic<long long> g_launches{0};

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

static std::atomic<int> g_pdl{-1};
bool pdl_enabled() {
  int v = g_pdl.load(std::memory_order_relaxed);
  if (v < 0) {
    const char* e = getenv("BSIG_PDL");
    v = (e != nullptr && atoi(e) == 0) ? 0 : 1;
    g_pdl.store(v, std::memory_order_relaxed);
  }
  return v != 0;
}

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

}  // namespace bsig

extern "C" const char* bsig_last_error(void) { return bsig::g_err; }
extern "C" int bsig_version(void) { return BSIG_VERSION; }
extern "C" int64_t bsig_launch_count(void) { return (int64_t)bsig::g_launches.load(); }
extern "C" int bsig_set_pdl(int enabled) {
  bsig::g_pdl.store(enabled ? 1 : 0);
  return 0;
}

extern "C" int bsig_device_info(int* sm, int* cc_major, int* cc_minor) {
  int dev = 0;
  BSIG_CUDA(cudaGetDevice(&dev));
  BSIG_CUDA(cudaDeviceGetAttribute(sm, cudaDevAttrMultiProcessorCount, dev));
  BSIG_CUDA(cudaDeviceGetAttribute(cc_major, cudaDevAttrComputeCapabilityMajor, dev));
  BSIG_CUDA(cudaDeviceGetAttribute(cc_minor, cudaDevAttrComputeCapabilityMinor, dev));
  return 0;
}
