// Microbenchmark: the fused kernel's exact UMMA operand pattern in a tight, fully unrolled loop (no barriers, no ring waits):
// per K16 step  A_hi*W_hi (collector fill), A_hi*W_lo (lastuse), A_lo*W_hi (A from TMEM for K-blocks 0..3, smem for 4..5).
#include <cstdio>
#include <cuda_runtime.h>
#include "sm100_ptx.cuh"
using namespace gnrf::ptx;

template <int N, int MODE>
__global__ void __launch_bounds__(128, 1) ub(long long* out) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint32_t tmem_ptr;
  __shared__ __align__(8) unsigned long long bar_mem;
  const uint32_t bar = smem_u32(&bar_mem);
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 57000; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw)[i] = 0x3c003c00u;
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
    if (elect_one()) {
      long long t0 = clock64();
      for (int rep = 0; rep < 4; ++rep) {
#pragma unroll
        for (int k16 = 0; k16 < 24; ++k16) {
          const int kb = k16 >> 2;
          const uint64_t a_hi = umma_desc_sw128(sb + kb * 16384 + (k16 & 3) * 32);
          const uint64_t a_lo = umma_desc_sw128(sb + 98304 + (kb & 1) * 16384 + (k16 & 3) * 32);
          const uint32_t slot = sb + 131072 + ((k16 >> 2) % 3) * 32768;
          const uint64_t b_hi = umma_desc_nosw(slot + (k16 & 3) * N * 32);
          const uint64_t b_lo = umma_desc_nosw(slot + (4 + (k16 & 3)) * N * 32 % 32768);
          if (MODE == 0) {          // kernel pattern
            umma_ss_a_fill(tm, a_hi, b_hi, idesc, 1u);
            umma_ss_a_lastuse(tm, a_hi, b_lo, idesc, 1u);
            if (kb < 4) umma_ts(tm, tm + 384 + kb * 32 + (k16 & 3) * 8, b_hi, idesc, 1u);
            else umma_ss(tm, a_lo, b_hi, idesc, 1u);
          } else if (MODE == 1) {   // plain SS, no collector hints
            umma_ss(tm, a_hi, b_hi, idesc, 1u);
            umma_ss(tm, a_hi, b_lo, idesc, 1u);
            umma_ss(tm, a_lo, b_hi, idesc, 1u);
          } else if (MODE == 2) {   // all TS
            umma_ts(tm, tm + 384 + (kb & 3) * 32 + (k16 & 3) * 8, b_hi, idesc, 1u);
            umma_ts(tm, tm + 384 + (kb & 3) * 32 + (k16 & 3) * 8, b_lo, idesc, 1u);
            umma_ts(tm, tm + 384 + (kb & 3) * 32 + (k16 & 3) * 8, b_hi, idesc, 1u);
          }
        }
      }
      umma_commit(bar);
      long long t1 = clock64();
      mbar_wait(bar, 0);
      long long t2 = clock64();
      if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    __syncwarp();
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) { tc_fence_after_sync(); tmem_dealloc_512(tmem); }
}

template <int N, int MODE>
void run(const char* name, long long* d) {
  cudaFuncSetAttribute(ub<N, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 230400);
  ub<N, MODE><<<148, 128, 230400>>>(d);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  long long h[2] = {0, 0};
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("%-40s %.1f cyc/MMA issued, %.1f cyc/MMA completed (ideal %d) %s\n", name, h[0] / 288.0, h[1] / 288.0, N / 2,
         e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  long long* d;
  cudaMalloc(&d, 64);
  run<128, 0>("N=128 fill/lastuse/TS(+SS kb>=4)", d);
  run<128, 1>("N=128 plain SS x3", d);
  run<128, 2>("N=128 all TS", d);
  run<256, 0>("N=256 fill/lastuse/TS(+SS kb>=4)", d);
  run<256, 1>("N=256 plain SS x3", d);
  run<192, 0>("N=192 fill/lastuse/TS(+SS kb>=4)", d);
  run<192, 1>("N=192 plain SS x3", d);
  return 0;
}
