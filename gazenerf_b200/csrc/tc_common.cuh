// Device helpers shared by the tcgen05 kernels (mlp_tc.cu, conv_tc.cu): UMMA-canonical SWIZZLE_128B K-major A tiles.
#pragma once
#include "sm100_ptx.cuh"

namespace gnrf {
namespace tc {

using namespace ptx;

// A tile = [128 rows x 64 bf16] (16 KB): 8-row groups of 1024 B, rows of 128 B, 16-byte chunk index XOR (row & 7).
__device__ __forceinline__ uint32_t a_row_offset(int row) { return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128); }

// x = hi + lo split of 64 consecutive K values of one row, packed as bf16 pairs.
__device__ __forceinline__ void split_row64(const float (&v)[64], uint32_t (&hi)[32], uint32_t (&lo)[32]) {
#pragma unroll
  for (int q = 0; q < 32; ++q) split2(v[2 * q], v[2 * q + 1], hi[q], lo[q]);
}
__device__ __forceinline__ void st_shared_row128(uint32_t addr_row, uint32_t sw, const uint32_t (&w)[32]) {
#pragma unroll
  for (int j = 0; j < 8; ++j)
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr_row + (((uint32_t)j ^ sw) << 4)), "r"(w[4 * j]),
                 "r"(w[4 * j + 1]), "r"(w[4 * j + 2]), "r"(w[4 * j + 3])
                 : "memory");
}

}  // namespace tc
}  // namespace gnrf
