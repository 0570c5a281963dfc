// Host API of the tcgen05 1x1-conv GEMM (conv_tc.cu), used by the neural-renderer orchestration.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>

namespace gnrf {
namespace tc {

enum ConvEpilogue { CONV_EPI_LRELU = 0, CONV_EPI_PSU = 1, CONV_EPI_LINEAR = 2, CONV_EPI_RELU = 3 };  // LINEAR: bias only

struct ConvLayerPlan {
  int N, K;            // output / input channels
  int n_chunks, chunk_n, k16_steps, n_kb;
  size_t stream_bytes, bias_floats, total_bytes;
};

ConvLayerPlan conv_layer_plan(int N, int K);
int conv_tc_pack(const ConvLayerPlan& pl, const float* W, const float* b, unsigned char* dst, cudaStream_t st);
// W(n,k) = W[n * sn + k * sk]; b may be null (zero bias)
int conv_tc_pack_strided(const ConvLayerPlan& pl, const float* W, long long sn, long long sk, const float* b, unsigned char* dst,
                         cudaStream_t st);
// out = epi(W X + b): X [n_img][K][HW] fp32 NCHW; mode LRELU -> out [n_img][N][HW]; mode PSU -> + res[n % Cres], pixel-shuffled
// to [n_img][N/4][2H][2W].
int conv_tc_launch(const ConvLayerPlan& pl, const unsigned char* packed, const float* X, float* out, const float* res, int Cres,
                   int n_img, int HW, int Wd, int mode, cudaStream_t st);

// Generic-GEMM options of the training path: image strides (elements), per-image bias, and an epilogue
//   v = act(acc + bias); v *= (mask > 0 ? 1 : mask_slope) for rows < mask_rows; v += add for rows < add_rows   (not with PSU).
struct ConvExtras {
  long long x_img_stride, out_img_stride;
  const float* bias_img;
  const float* mask; long long mask_img_stride; int mask_rows; float mask_slope;
  const float* add; long long add_img_stride; int add_rows;
};
int conv_tc_launch_ex(const ConvLayerPlan& pl, const unsigned char* packed, const float* X, float* out, const float* res, int Cres,
                      int n_img, int HW, int Wd, int mode, const ConvExtras& ex, cudaStream_t st);

}  // namespace tc
}  // namespace gnrf
