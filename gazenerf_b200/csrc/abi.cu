// libgnrf: error plumbing + device check (see include/gnrf.h).
#include "common.cuh"

#include <atomic>

namespace gnrf {

char* error_buffer() {
  static thread_local char buf[512] = {0};
  return buf;
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(error_buffer(), 512, fmt, ap);
  va_end(ap);
  return code;
}

static std::atomic<unsigned long long> g_launches{0};
void count_launches(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

}  // namespace gnrf

extern "C" unsigned long long gnrf_launch_count(void) { return gnrf::g_launches.load(std::memory_order_relaxed); }

extern "C" int gnrf_abi_version(void) { return GNRF_ABI_VERSION; }

extern "C" const char* gnrf_last_error(void) { return gnrf::error_buffer(); }

extern "C" int gnrf_device_check(void) {
  int dev = 0;
  GNRF_CUDA(cudaGetDevice(&dev));
  int major = 0, minor = 0;
  GNRF_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  GNRF_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  if (major != 10 || minor != 0)
    return gnrf::fail(GNRF_ERR_UNSUPPORTED, "gnrf_device_check: device %d is sm_%d%d, libgnrf is built for sm_100a only", dev, major, minor);
  return GNRF_OK;
}
