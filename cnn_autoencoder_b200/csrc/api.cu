// Library-level entry points of the C ABI: error text, device probe, launch counter.
#include <atomic>
#include <stdarg.h>

#include "cae_common.cuh"

namespace {
thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};
}  // namespace

void cae_set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void cae_count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

extern "C" int cae_abi_version(void) { return CAE_ABI_VERSION; }

extern "C" const char *cae_last_error(void) { return g_err; }

extern "C" uint64_t cae_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" int cae_device_info(int *sm_count, int *cc_major, int *cc_minor) {
  int dev = 0, sms = 0, major = 0, minor = 0;
  CAE_CUDA(cudaGetDevice(&dev));
  CAE_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  CAE_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  CAE_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  if (sm_count) *sm_count = sms;
  if (cc_major) *cc_major = major;
  if (cc_minor) *cc_minor = minor;
  CAE_CHECK(major == 10, 7, "device compute capability %d.%d: this library carries sm_100a code only",
            major, minor);
  return 0;
}
