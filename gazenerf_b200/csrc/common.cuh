// Shared host/device helpers for libgnrf (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdio>
#include <cstdint>

#include "gnrf.h"

namespace gnrf {

// thread-local error text returned by gnrf_last_error()
char* error_buffer();
int fail(int code, const char* fmt, ...);
void count_launches(int n);  // feeds gnrf_launch_count()

#define GNRF_CHECK_ARG(cond)                                                             \
  do {                                                                                   \
    if (!(cond)) return ::gnrf::fail(GNRF_ERR_ARG, "%s: argument check failed: %s", __func__, #cond); \
  } while (0)

#define GNRF_CUDA(call)                                                                  \
  do {                                                                                   \
    cudaError_t e__ = (call);                                                            \
    if (e__ != cudaSuccess)                                                              \
      return ::gnrf::fail(GNRF_ERR_CUDA, "%s: %s -> %s", __func__, #call, cudaGetErrorString(e__)); \
  } while (0)

#define GNRF_LAUNCH_CHECK()                                                              \
  do {                                                                                   \
    cudaError_t e__ = cudaGetLastError();                                                \
    if (e__ != cudaSuccess)                                                              \
      return ::gnrf::fail(GNRF_ERR_CUDA, "%s: kernel launch failed: %s", __func__, cudaGetErrorString(e__)); \
  } while (0)

// Per-device one-time initialisation (cudaFuncSetAttribute opt-ins are per device / context): runs `init` the first time `key` is
// seen on the CURRENT device, under a process-wide mutex so that concurrent first calls from several threads are safe.
// Returns the current device's SM count through *n_sm.  key < 32.
int device_once_begin(int key, int* dev, int* n_sm, bool* need_init);   // takes the lock when *need_init
void device_once_end(int key, int dev, bool ok);                        // releases it
template <class F>
static inline int device_once(int key, int* n_sm, F&& init) {
  int dev = 0;
  bool need = false;
  int rc = device_once_begin(key, &dev, n_sm, &need);
  if (rc != GNRF_OK || !need) return rc;
  rc = init();
  device_once_end(key, dev, rc == GNRF_OK);
  return rc;
}
enum { kOnceMlpTc = 0, kOnceMlpSimt = 1, kOnceConvTc = 2, kOnceWgradTc = 3, kOnceNrFused = 4, kOnceMlpTrain = 5, kOnceLinHl2 = 6,
       kOnceLinHl1 = 7, kOnceWgHl2 = 8, kOnceWgHl1 = 9, kOnceLinHlSm = 10 };

static inline cudaStream_t as_stream(gnrf_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

}  // namespace gnrf
