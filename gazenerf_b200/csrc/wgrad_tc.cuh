// Host API of the tcgen05 weight-gradient GEMM (wgrad_tc.cu).
#pragma once
#include <cuda_runtime.h>
#include <cstddef>

namespace gnrf {
namespace tc {

struct WgradPlan {
  int rows_x, n_mt, n_ch, chunk_n, n_kb_total, n_split, n_items;
  size_t partial_bytes;
};

WgradPlan wgrad_plan(int N_dy, int K_x, int n_img, int HW, bool want_bias);
// dW [N_dy][K_x] (=, or += when accumulate) sum_img sum_p dY[img][n][p] X[img][k][p];
// db (nullable): db_sum == 0 -> [n_img][N_dy] per-image sum_p dY; db_sum != 0 -> [N_dy] summed over the images (honours accumulate).
int wgrad_tc_launch(const float* dY, long long dy_img_stride, const float* X, long long x_img_stride, int N_dy, int K_x, int n_img,
                    int HW, float* dW, float* db_img, int db_sum, int accumulate, void* ws, size_t ws_bytes, cudaStream_t st);

}  // namespace tc
}  // namespace gnrf
