// Library-level entry points of the C ABI: error text, device probe, launch counter.
#include <atomic>
#include <mutex>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

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

namespace {
const char *g_knob_val[CAE_KNOB_COUNT];
std::once_flag g_knob_once;

void read_knobs() {
  static const char *names[CAE_KNOB_COUNT] = {
#define X(n) "CAE_" #n,
      CAE_KNOB_LIST(X)
#undef X
  };
  const char *dbg = getenv("CAE_DEBUG");
  const bool on = dbg && strcmp(dbg, "1") == 0;
  for (int i = 0; i < CAE_KNOB_COUNT; ++i) {
    const char *v = getenv(names[i]);
    g_knob_val[i] = nullptr;
    if (!v) continue;
    if (on) g_knob_val[i] = strdup(v);
    else fprintf(stderr, "cae_b200: ignoring bring-up knob %s (set CAE_DEBUG=1 to enable it)\n", names[i]);
  }
}
}  // namespace

const char *cae_knob(CaeKnob k) {
  std::call_once(g_knob_once, read_knobs);
  return g_knob_val[k];
}

int cae_sm_count() {
  static std::atomic<int> cached[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  int v = cached[dev].load(std::memory_order_relaxed);
  if (v > 0) return v;
  if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
  cached[dev].store(v, std::memory_order_relaxed);
  return v;
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
