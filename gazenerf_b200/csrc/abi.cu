// libgnrf: error plumbing + device check (see include/gnrf.h).
#include "common.cuh"

#include <atomic>
#include <mutex>

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

// ---- per-device one-time initialisation (see common.cuh) ------------------------------------------------------------
namespace {
constexpr int kMaxDevices = 64;
std::mutex g_once_mutex;
uint32_t g_once_done[kMaxDevices] = {0};   // bit `key` set: initialised on that device     (guarded by g_once_mutex)
int g_sm_count[kMaxDevices] = {0};
}  // namespace

int device_once_begin(int key, int* dev, int* n_sm, bool* need_init) {
  *need_init = false;
  GNRF_CUDA(cudaGetDevice(dev));
  if (*dev < 0 || *dev >= kMaxDevices) return fail(GNRF_ERR_UNSUPPORTED, "device ordinal %d out of range", *dev);
  g_once_mutex.lock();
  if (g_sm_count[*dev] == 0) {
    int n = 0;
    cudaError_t e = cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, *dev);
    if (e != cudaSuccess) {
      g_once_mutex.unlock();
      return fail(GNRF_ERR_CUDA, "cudaDeviceGetAttribute(MultiProcessorCount) -> %s", cudaGetErrorString(e));
    }
    g_sm_count[*dev] = n;
  }
  *n_sm = g_sm_count[*dev];
  if (g_once_done[*dev] & (1u << key)) {
    g_once_mutex.unlock();
    return GNRF_OK;
  }
  *need_init = true;   // the caller runs its init and then calls device_once_end, which unlocks
  return GNRF_OK;
}

void device_once_end(int key, int dev, bool ok) {
  if (ok) g_once_done[dev] |= (1u << key);
  g_once_mutex.unlock();
}

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
