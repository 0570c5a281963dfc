// Microbenchmark (not part of the library): tcgen05.mma (kind::f16, M=128, K=16) rate with the A operand in the no-swizzle K-major
// core-matrix layout (what nr_fused.cuh uses for t1 / A3 tiles) vs the SWIZZLE_128B layout, in the bf16x3 pattern
// (A_hi*B_hi, A_lo*B_hi, A_hi*B_lo per K16 step, new A and B addresses every step), for the N values of the fused neural renderer.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I gazenerf_b200/csrc -I include -o tests/ubench/ub_alayout tests/ubench/ub_alayout.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "sm100_ptx.cuh"
using namespace gnrf::ptx;

template <int N, int A_NOSW, int NSTEP>
__global__ void __launch_bounds__(128, 1) ub(long long* out) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint32_t tmem_ptr;
  __shared__ __align__(8) unsigned long long bar_mem;
  const uint32_t bar = smem_u32(&bar_mem);
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 49152; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw)[i] = 0x3c003c00u;
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
    long long t0 = 0, t2 = 0;
    uint32_t phase = 0;
    for (int rep = 0; rep < 4; ++rep) {
      __syncwarp();
      t0 = clock64();
      if (elect_one()) {
#pragma unroll 4
        for (int i = 0; i < NSTEP; ++i) {
          // A region: 8 K16 steps x (hi 4 KB | lo 4 KB) = 64 KB at `sb`; B region at sb + 64 KB: 8 slices x (hi | lo) x N*32 B
          const int s = i & 7;
          uint64_t a_hi, a_lo;
          if (A_NOSW) {
            a_hi = umma_desc_nosw(sb + s * 8192);
            a_lo = umma_desc_nosw(sb + s * 8192 + 4096);
          } else {   // SW128: K-block (s >> 2) of 16 KB hi at sb, lo at sb + 32 KB; +32 B per K16 step inside the block
            a_hi = umma_desc_sw128(sb + (s >> 2) * 16384 + (s & 3) * 32);
            a_lo = umma_desc_sw128(sb + 32768 + (s >> 2) * 16384 + (s & 3) * 32);
          }
          const uint64_t b_hi = umma_desc_nosw(sb + 65536 + s * (2 * N * 32));
          const uint64_t b_lo = umma_desc_nosw(sb + 65536 + s * (2 * N * 32) + N * 32);
          umma_ss(tm, a_hi, b_hi, idesc, 1u);
          umma_ss(tm, a_lo, b_hi, idesc, 1u);
          umma_ss(tm, a_hi, b_lo, idesc, 1u);
        }
        umma_commit(bar);
      }
      __syncwarp();
      mbar_wait(bar, phase);
      phase ^= 1;
      t2 = clock64();
    }
    if (threadIdx.x == 32 && blockIdx.x == 0) { out[0] = t2 - t0; }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) { tc_fence_after_sync(); tmem_dealloc_512(tmem); }
}

template <int N, int A_NOSW>
void run(long long* d) {
  constexpr int NS = 128;
  cudaFuncSetAttribute(ub<N, A_NOSW, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  ub<N, A_NOSW, NS><<<148, 128, 200 * 1024>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2] = {0, 0};
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("N=%3d A=%-6s : %7.1f cycles per UMMA (ideal %d) %s\n", N, A_NOSW ? "nosw" : "sw128", (double)h[0] / (3.0 * NS), N / 2,
         e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  long long* d;
  cudaMalloc(&d, 64);
  run<32, 0>(d); run<32, 1>(d);
  run<64, 0>(d); run<64, 1>(d);
  run<128, 0>(d); run<128, 1>(d);
  run<144, 0>(d); run<144, 1>(d);
  run<256, 0>(d); run<256, 1>(d);
  return 0;
}
