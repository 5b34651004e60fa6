// Status, error string and launch accounting for the C ABI (include/b200lrcn.h).
#include "common.cuh"

#include <atomic>
#include <stdarg.h>
#include <string.h>

namespace {
thread_local char g_err[512] = "";
std::atomic<long> g_launches{0};
int g_sms = 0;
}  // namespace

void b2_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void b2_count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int b2_num_sms() {
  if (g_sms == 0) {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && sms > 0)
      g_sms = sms;
    else
      g_sms = 148;
  }
  return g_sms;
}

B2_API int b2_abi_version(void) { return 1; }
B2_API const char* b2_last_error(void) { return g_err; }
B2_API long b2_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
// replaying a captured CUDA graph launches its kernels without passing through the entry points: the host layer
// credits them here so that b2_launch_count() stays the number of kernels executed
B2_API int b2_add_launch_count(long n) {
  g_launches.fetch_add(n, std::memory_order_relaxed);
  return 0;
}

// 0 when the current device is a compute-capability 10.x part (B200); otherwise an error.
// There is no CPU or other-architecture fallback anywhere in this library.
B2_API int b2_device_check(void) {
  int dev = 0, major = 0, minor = 0;
  B2_CUDA_CHECK(cudaGetDevice(&dev));
  B2_CUDA_CHECK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  B2_CUDA_CHECK(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  B2_ARG_CHECK(major == 10, "b200lrcn needs an sm_100a device (B200); found sm_%d%d", major, minor);
  return 0;
}
