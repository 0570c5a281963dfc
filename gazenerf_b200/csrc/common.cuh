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

static inline cudaStream_t as_stream(gnrf_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

}  // namespace gnrf
