// Adjoint operators of the neural renderer's fixed linear stages (train_ops.cu), used by the NR backward orchestration.
#pragma once
#include <cuda_runtime.h>

namespace gnrf {
// g_in = Blur^T (g_out * (act > 0 ? 1 : slope));  act may be null.  [planes][H][W]
int launch_blur_adj(const float* g_out, const float* act, float slope, int planes, int H, int Wd, float* g_in, cudaStream_t st);
// g_in [planes][H][W] = BilinearUp2^T g_out [planes][2H][2W]
int launch_up2_adj(const float* g_out, int planes, int H, int Wd, float* g_in, cudaStream_t st);
int launch_sigmoid_bwd(const float* g_img, const float* img, long long total, float* g, cudaStream_t st);
// backward of sh = pixel_shuffle2(LReLU(v) + repeat(x,4)):  g_pre [N][4ci][H][W], g_res [N][ci][H][W]
int launch_psu_bwd(const float* g_sh, const float* sh, const float* x, int N, int ci, int H, int Wd, float* g_pre, float* g_res,
                   cudaStream_t st);
}  // namespace gnrf
