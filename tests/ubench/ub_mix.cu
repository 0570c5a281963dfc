// Microbenchmark: N=128 / N=256 UMMA bursts with the fused kernel's operand pattern (A_hi fill, A_hi lastuse, A_lo from TMEM),
// optionally with a concurrent TMA stream into the smem ring (real bulk copies from global, L2 resident) and/or a warpgroup doing
// epilogue-like work (tcgen05.ld + st.shared + tcgen05.st).  Reports cycles per stage of 384 ideal tensor cycles.
#include <cstdio>
#include <cuda_runtime.h>
#include "sm100_ptx.cuh"
using namespace gnrf::ptx;
constexpr int kSlots = 6;

template <int N, int BURST3, int PATTERN, int TMA, int EPI>
__global__ void __launch_bounds__(192, 1) ub(long long* out, int n_stages, const unsigned char* gsrc) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint32_t tmem_ptr;
  __shared__ __align__(8) unsigned long long bars[2 * kSlots + 1];
  __shared__ volatile int stop_flag;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 50000; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2 * kSlots + 1; ++i) mbar_init(smem_u32(&bars[i]), 1);
    fence_mbar_init();
    stop_flag = 0;
  }
  if (warp == 5) tmem_alloc_512(smem_u32(&tmem_ptr));
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_ptr;
  auto full = [&](int s) { return smem_u32(&bars[s]); };
  auto empty = [&](int s) { return smem_u32(&bars[kSlots + s]); };
  const uint32_t ring = base + 131072;
  if (warp == 4) {
    if (elect_one()) {
      uint32_t slot = 0, phase = 0;
      const unsigned char* src = gsrc;
      for (int s = 0; s < n_stages; ++s) {
        mbar_wait_spin(empty(slot), phase ^ 1);
        if (TMA) {
          mbar_arrive_expect_tx(full(slot), 16384);
          bulk_g2s(ring + slot * 16384, src, 16384, full(slot));
          src += 16384;
          if (src >= gsrc + (4u << 20)) src = gsrc;
        } else {
          mbar_arrive(full(slot));
        }
        if (++slot == kSlots) { slot = 0; phase ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 5) {
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
    const uint32_t sb = __shfl_sync(0xffffffffu, base, 0);
    long long t0 = 0, t1 = 0;
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, N);
      uint32_t slot = 0, phase = 0;
      t0 = clock64();
      for (int s = 0; s < n_stages; ++s) {
        mbar_wait_spin(full(slot), phase);
        tc_fence_after_sync();
        const uint32_t b0 = ring + slot * 16384;
#pragma unroll
        for (int i = 0; i < BURST3; ++i) {
          const uint64_t a = umma_desc_sw128(sb + ((s * BURST3 + i) % 24 / 4) * 16384 + (i & 3) * 32);
          const uint64_t bh = umma_desc_nosw(b0 + i * N * 32), bl = umma_desc_nosw(b0 + (BURST3 + i) * N * 32);
          if (PATTERN == 0) {   // plain SS x3
            umma_ss(tm, a, bh, idesc, 1u); umma_ss(tm, a, bl, idesc, 1u); umma_ss(tm, a, bh, idesc, 1u);
          } else {              // kernel pattern: fill, lastuse, TS
            umma_ss_a_fill(tm, a, bh, idesc, 1u); umma_ss_a_lastuse(tm, a, bl, idesc, 1u);
            umma_ts(tm, tm + 384 + (i & 3) * 8, bh, idesc, 1u);
          }
        }
        umma_commit(empty(slot));
        if (++slot == kSlots) { slot = 0; phase ^= 1; }
      }
      t1 = clock64();
      stop_flag = 1;
      if (blockIdx.x == 0) out[0] = t1 - t0;
    }
    __syncwarp();
  } else if (EPI) {
    // epilogue-like background load on warps 0-3: TMEM loads, conversions, swizzled smem stores, TMEM stores
    const uint32_t t_lane = tmem + ((uint32_t)(warp * 32) << 16);
    const int row = threadIdx.x;
    int it = 0;
    while (!stop_flag) {
      uint32_t r0[32], r1[32];
      tmem_ld32(t_lane + (it % 6) * 64, r0);
      tmem_ld32(t_lane + (it % 6) * 64 + 32, r1);
      tmem_wait_ld();
      uint32_t hi[32], lo[32];
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        split2(__uint_as_float(r0[2 * q]) + 1.0f, __uint_as_float(r0[2 * q + 1]), hi[q], lo[q]);
        split2(__uint_as_float(r1[2 * q]) + 1.0f, __uint_as_float(r1[2 * q + 1]), hi[16 + q], lo[16 + q]);
      }
      const uint32_t addr = base + 98304 + (row >> 3) * 1024 + (row & 7) * 128;   // a scratch A tile (not read by the MMAs)
#pragma unroll
      for (int j = 0; j < 8; ++j)
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr + ((j ^ (row & 7)) << 4)), "r"(hi[4 * j]), "r"(hi[4 * j + 1]),
                     "r"(hi[4 * j + 2]), "r"(hi[4 * j + 3]) : "memory");
      tmem_st32(t_lane + 480, lo);
      tmem_wait_st();
      fence_proxy_async_smem();
      ++it;
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 5) { tc_fence_after_sync(); tmem_dealloc_512(tmem); }
}

template <int N, int B3, int P, int T, int E>
void run(const char* name, long long* d, const unsigned char* g) {
  const int n_stages = 600;
  cudaFuncSetAttribute(ub<N, B3, P, T, E>, cudaFuncAttributeMaxDynamicSharedMemorySize, 230400);
  ub<N, B3, P, T, E><<<148, 192, 230400>>>(d, n_stages, g);
  cudaError_t e = cudaGetLastError(); if (e == cudaSuccess) e = cudaDeviceSynchronize();
  long long h = 0;
  cudaMemcpy(&h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("%-58s %.0f cyc/stage (ideal 384)  %s\n", name, h / (double)n_stages, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  long long* d;
  unsigned char* g;
  cudaMalloc(&d, 64);
  cudaMalloc(&g, 8u << 20);
  cudaMemset(g, 0x3c, 8u << 20);
  run<128, 2, 0, 0, 0>("N=128 SSx3                 no TMA, no epi", d, g);
  run<128, 2, 1, 0, 0>("N=128 fill/lastuse/TS      no TMA, no epi", d, g);
  run<128, 2, 1, 1, 0>("N=128 fill/lastuse/TS      TMA          ", d, g);
  run<128, 2, 1, 0, 1>("N=128 fill/lastuse/TS      epi          ", d, g);
  run<128, 2, 1, 1, 1>("N=128 fill/lastuse/TS      TMA + epi    ", d, g);
  run<128, 2, 0, 1, 1>("N=128 SSx3                 TMA + epi    ", d, g);
  run<256, 1, 1, 0, 0>("N=256 fill/lastuse/TS      no TMA, no epi", d, g);
  run<256, 1, 1, 1, 0>("N=256 fill/lastuse/TS      TMA          ", d, g);
  run<256, 1, 1, 1, 1>("N=256 fill/lastuse/TS      TMA + epi    ", d, g);
  run<192, 1, 1, 0, 0>("N=192 x3 fill/lastuse/TS   no TMA, no epi (ideal 288)", d, g);
  run<192, 1, 1, 1, 1>("N=192 x3 fill/lastuse/TS   TMA + epi      (ideal 288)", d, g);
  run<128, 1, 1, 1, 1>("N=128 x3 fill/lastuse/TS   TMA + epi      (ideal 192)", d, g);
  run<256, 2, 1, 1, 1>("N=256 x6 fill/lastuse/TS   TMA + epi      (ideal 768)", d, g);
  run<128, 4, 1, 1, 1>("N=128 x12 fill/lastuse/TS  TMA + epi      (ideal 768)", d, g);
  run<128, 4, 1, 0, 0>("N=128 x12 fill/lastuse/TS  no TMA no epi  (ideal 768)", d, g);
  return 0;
}
