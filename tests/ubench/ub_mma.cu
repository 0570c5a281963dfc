// Microbenchmark (not part of the library): issue/execute rate of tcgen05.mma (kind::f16, M=128, K=16) for several N,
// A from smem (SS) or TMEM (TS), measured with clock64 around a batch of back-to-back MMAs + one commit.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I gazenerf_b200/csrc -I include -o /tmp/ub_mma tests/ubench/ub_mma.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "sm100_ptx.cuh"
using namespace gnrf::ptx;

template <int N, int TS, int NMMA, int DISTINCT_B>
__global__ void __launch_bounds__(128, 1) ub(long long* out) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint32_t tmem_ptr;
  __shared__ __align__(8) unsigned long long bar_mem;
  const uint32_t bar = smem_u32(&bar_mem);
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 40960; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc_512(smem_u32(&tmem_ptr));
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_ptr;
  if (warp == 1) {
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
    const uint32_t sb = __shfl_sync(0xffffffffu, base, 0);
    constexpr uint32_t idesc = umma_idesc_bf16(128, N);
    long long t0 = 0, t1 = 0, t2 = 0;
    uint32_t phase = 0;
    for (int rep = 0; rep < 4; ++rep) {
      __syncwarp();
      t0 = clock64();
      if (elect_one()) {
#pragma unroll 8
        for (int i = 0; i < NMMA; ++i) {
          const uint64_t a = umma_desc_sw128(sb + (i & 3) * 32);
          const uint64_t b = umma_desc_nosw(sb + 65536 + (DISTINCT_B ? (i & 7) * N * 32 : 0));
          if (TS) umma_ts(tm, tm + 384 + (i & 3) * 8, b, idesc, 1u);
          else umma_ss(tm, a, b, idesc, 1u);
        }
        umma_commit(bar);
      }
      __syncwarp();
      t1 = clock64();
      mbar_wait(bar, phase);
      phase ^= 1;
      t2 = clock64();
    }
    if (threadIdx.x == 32 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) { tc_fence_after_sync(); tmem_dealloc_512(tmem); }
}

template <int N, int TS, int DB>
void run(const char* name, long long* d) {
  constexpr int NM = 256;
  cudaFuncSetAttribute(ub<N, TS, NM, DB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  ub<N, TS, NM, DB><<<148, 128, 200 * 1024>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("%-28s issue %.1f cyc/MMA   issue+complete %.1f cyc/MMA   (ideal exec %d)  %s\n", name, h[0] / (double)NM, h[1] / (double)NM,
         N / 2, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  long long* d;
  cudaMalloc(&d, 64);
  run<64, 0, 0>("SS N=64  sameB", d);
  run<128, 0, 0>("SS N=128 sameB", d);
  run<128, 0, 1>("SS N=128 distinctB", d);
  run<256, 0, 0>("SS N=256 sameB", d);
  run<256, 0, 1>("SS N=256 distinctB", d);
  run<128, 1, 1>("TS N=128 distinctB", d);
  run<256, 1, 1>("TS N=256 distinctB", d);
  run<192, 0, 1>("SS N=192 distinctB", d);
  return 0;
}
