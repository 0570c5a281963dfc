// Microbenchmark: cost of the per-stage protocol around a burst of UMMAs (wait on an always-ready mbarrier ring, fence, elect,
// BURST x N-wide UMMAs, commit), for several variants of the loop.  clock64 per stage, averaged.
#include <cstdio>
#include <cuda_runtime.h>
#include "sm100_ptx.cuh"
using namespace gnrf::ptx;

constexpr int kSlots = 6;

template <int N, int BURST, int VARIANT>
__global__ void __launch_bounds__(192, 1) ub(long long* out, int n_stages) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint32_t tmem_ptr;
  __shared__ __align__(8) unsigned long long bars[2 * kSlots + 1];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 40960; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2 * kSlots + 1; ++i) mbar_init(smem_u32(&bars[i]), 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc_512(smem_u32(&tmem_ptr));
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_ptr;
  auto full = [&](int s) { return smem_u32(&bars[s]); };
  auto empty = [&](int s) { return smem_u32(&bars[kSlots + s]); };
  if (warp == 2) {
    // fake producer: re-arms "full" whenever the consumer frees a slot (no data movement)
    if (lane == 0) {
      uint32_t slot = 0, phase = 0;
      for (int s = 0; s < n_stages; ++s) {
        mbar_wait(empty(slot), phase ^ 1);
        mbar_arrive(full(slot));
        if (++slot == kSlots) { slot = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
    const uint32_t sb = __shfl_sync(0xffffffffu, base, 0);
    constexpr uint32_t idesc = umma_idesc_bf16(128, N);
    uint32_t slot = 0, phase = 0;
    long long t0 = clock64();
    if (VARIANT == 0) {
      // as in mlp_tc_kernel: wait -> fence -> elect { burst; commit } -> syncwarp
      for (int s = 0; s < n_stages; ++s) {
        mbar_wait(full(slot), phase);
        tc_fence_after_sync();
        const uint32_t b0 = sb + 65536 + slot * 16384;
        if (elect_one()) {
#pragma unroll
          for (int i = 0; i < BURST; ++i) umma_ss(tm, umma_desc_sw128(sb + (i & 3) * 32), umma_desc_nosw(b0 + (i & 1) * N * 32), idesc, 1u);
          umma_commit(empty(slot));
        }
        __syncwarp();
        if (++slot == kSlots) { slot = 0; phase ^= 1; }
      }
    } else if (VARIANT == 1) {
      // single elected lane runs the whole loop (no per-stage elect / syncwarp), bare try_wait spin without the clock-based timeout
      if (elect_one()) {
        for (int s = 0; s < n_stages; ++s) {
          while (!mbar_try_wait(full(slot), phase)) {}
          tc_fence_after_sync();
          const uint32_t b0 = sb + 65536 + slot * 16384;
#pragma unroll
          for (int i = 0; i < BURST; ++i) umma_ss(tm, umma_desc_sw128(sb + (i & 3) * 32), umma_desc_nosw(b0 + (i & 1) * N * 32), idesc, 1u);
          umma_commit(empty(slot));
          if (++slot == kSlots) { slot = 0; phase ^= 1; }
        }
      }
      __syncwarp();
    } else if (VARIANT == 2) {
      // converged warp, but the barrier of stage s+1 is probed BEFORE the burst of stage s (its latency overlaps the MMAs)
      bool ready = mbar_try_wait(full(slot), phase);
      for (int s = 0; s < n_stages; ++s) {
        if (!ready) mbar_wait(full(slot), phase);
        tc_fence_after_sync();
        const uint32_t b0 = sb + 65536 + slot * 16384;
        uint32_t nslot = slot + 1, nphase = phase;
        if (nslot == kSlots) { nslot = 0; nphase ^= 1; }
        if (elect_one()) {
#pragma unroll
          for (int i = 0; i < BURST; ++i) umma_ss(tm, umma_desc_sw128(sb + (i & 3) * 32), umma_desc_nosw(b0 + (i & 1) * N * 32), idesc, 1u);
          umma_commit(empty(slot));
        }
        __syncwarp();
        ready = (s + 1 < n_stages) ? mbar_try_wait(full(nslot), nphase) : true;
        slot = nslot; phase = nphase;
      }
    } else if (VARIANT == 3) {
      // no ring protocol at all: just bursts + commits (upper bound)
      if (elect_one()) {
        for (int s = 0; s < n_stages; ++s) {
          const uint32_t b0 = sb + 65536 + slot * 16384;
#pragma unroll
          for (int i = 0; i < BURST; ++i) umma_ss(tm, umma_desc_sw128(sb + (i & 3) * 32), umma_desc_nosw(b0 + (i & 1) * N * 32), idesc, 1u);
          umma_commit(empty(slot));
          if (++slot == kSlots) { slot = 0; phase ^= 1; }
        }
      }
      __syncwarp();
    }
    long long t1 = clock64();
    if (lane == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) { tc_fence_after_sync(); tmem_dealloc_512(tmem); }
}

template <int N, int BURST, int V>
void run(const char* name, long long* d) {
  const int n_stages = 600;
  cudaFuncSetAttribute(ub<N, BURST, V>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  ub<N, BURST, V><<<148, 192, 200 * 1024>>>(d, n_stages);
  cudaError_t e = cudaDeviceSynchronize();
  long long h = 0;
  cudaMemcpy(&h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("%-44s %.0f cyc/stage  (MMA exec %d)  %s\n", name, h / (double)n_stages, BURST * N / 2, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  long long* d;
  cudaMalloc(&d, 64);
  run<128, 6, 0>("V0 current loop        N=128 x6", d);
  run<256, 3, 0>("V0 current loop        N=256 x3", d);
  run<128, 6, 1>("V1 single-lane loop    N=128 x6", d);
  run<256, 3, 1>("V1 single-lane loop    N=256 x3", d);
  run<128, 6, 2>("V2 early probe         N=128 x6", d);
  run<256, 3, 2>("V2 early probe         N=256 x3", d);
  run<128, 6, 3>("V3 no ring (bound)     N=128 x6", d);
  run<256, 3, 3>("V3 no ring (bound)     N=256 x3", d);
  run<128, 12, 0>("V0 current loop        N=128 x12", d);
  run<128, 12, 2>("V2 early probe         N=128 x12", d);
  return 0;
}
