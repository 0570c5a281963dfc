// Fused data terms of GazeNeRFLoss (losses/gazenerf_loss.py:294-352 calc_data_loss, masks from calc_total_loss :420-424).
// The reference builds boolean masks, gathers res_img[mask] / gt[mask] (a device->host sync per gather for the output size) and calls
// l1_loss / mse_loss on the gathered vectors.  mean over a boolean gather == sum(mask * d) / count(mask): ONE streaming pass over the
// four rendered images + gt + the four mask planes produces the five loss terms, and one more pass writes the image gradients.
//   head  = face_mask >= .5 && full_eye < .5      -> l1|mse(merge_img,      gt)
//   face  = face_mask >= .5 && left < .5 && right < .5 -> l1|mse(merge_img_face, gt)
//   eyes  = left >= .5 || right >= .5             -> l1|mse(merge_img_eyes, gt)
//   nonhead = face_mask < .5                      -> mean((merge_img - bg_value)^2)
//   bg                                            -> mean((bg_img - bg_value)^2)
#include "common.cuh"

namespace gnrf {

constexpr int kLossTerms = 5;    // head, eyes, face, nonhead, bg
constexpr int kLossSums = 9;     // S_head N_head S_eyes N_eyes S_face N_face S_nonhead N_nonhead S_bg

struct LossArgs {
  const float* img_face; const float* img_eyes; const float* img; const float* bg_img; const float* gt;
  const float* face_mask; const float* full_eye; const float* left_eye; const float* right_eye;
  int B, HW, use_l1;
  float bg_value;
};

__device__ __forceinline__ void pixel_masks(const LossArgs& a, size_t bp, float& m_head, float& m_face, float& m_eyes, float& m_nh) {
  const float fm = a.face_mask[bp], fe = a.full_eye[bp], le = a.left_eye[bp], re = a.right_eye[bp];
  m_head = (fm >= 0.5f && fe < 0.5f) ? 1.0f : 0.0f;
  m_face = (fm >= 0.5f && le < 0.5f && re < 0.5f) ? 1.0f : 0.0f;
  m_eyes = (le >= 0.5f || re >= 0.5f) ? 1.0f : 0.0f;
  m_nh = fm < 0.5f ? 1.0f : 0.0f;
}

// partial[block][9]; one thread per (b, pixel), channels looped.
__global__ void __launch_bounds__(256) data_loss_partial_kernel(const LossArgs a, float* __restrict__ partial) {
  float s[kLossSums];
#pragma unroll
  for (int i = 0; i < kLossSums; ++i) s[i] = 0.0f;
  const long long total = (long long)a.B * a.HW;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(idx / a.HW), p = (int)(idx - (long long)b * a.HW);
    float m_head, m_face, m_eyes, m_nh;
    pixel_masks(a, (size_t)idx, m_head, m_face, m_eyes, m_nh);
    s[1] += m_head; s[3] += m_eyes; s[5] += m_face; s[7] += m_nh;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const size_t o = ((size_t)b * 3 + c) * a.HW + p;
      const float g = a.gt[o];
      const float d0 = a.img[o] - g, d1 = a.img_eyes[o] - g, d2 = a.img_face[o] - g;
      s[0] += m_head * (a.use_l1 ? fabsf(d0) : d0 * d0);
      s[2] += m_eyes * (a.use_l1 ? fabsf(d1) : d1 * d1);
      s[4] += m_face * (a.use_l1 ? fabsf(d2) : d2 * d2);
      const float tv = a.img[o] - a.bg_value;
      s[6] += m_nh * tv * tv;
      if (b == 0) {
        const float t = a.bg_img[(size_t)c * a.HW + p] - a.bg_value;
        s[8] += t * t;
      }
    }
  }
  __shared__ float red[8][kLossSums];
#pragma unroll
  for (int i = 0; i < kLossSums; ++i)
    for (int o = 16; o > 0; o >>= 1) s[i] += __shfl_xor_sync(0xffffffffu, s[i], o);
  if ((threadIdx.x & 31) == 0)
#pragma unroll
    for (int i = 0; i < kLossSums; ++i) red[threadIdx.x >> 5][i] = s[i];
  __syncthreads();
  if (threadIdx.x < kLossSums) {
    float t = 0.0f;
    for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
    partial[(size_t)blockIdx.x * kLossSums + threadIdx.x] = t;
  }
}

// sums[9] (fixed-order reduction over blocks) and terms[5] = (head, eyes, face, nonhead, bg); an empty mask gives 0/0 = NaN exactly like
// l1_loss on an empty gather in the reference.
__global__ void data_loss_final_kernel(const float* __restrict__ partial, int n_blocks, int HW, float* __restrict__ sums,
                                       float* __restrict__ terms) {
  __shared__ float s[kLossSums];
  if (threadIdx.x < kLossSums) {
    float t = 0.0f;
    for (int b = 0; b < n_blocks; ++b) t += partial[(size_t)b * kLossSums + threadIdx.x];
    s[threadIdx.x] = t;
    sums[threadIdx.x] = t;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    terms[0] = s[0] / (3.0f * s[1]);
    terms[1] = s[2] / (3.0f * s[3]);
    terms[2] = s[4] / (3.0f * s[5]);
    terms[3] = s[6] / (3.0f * s[7]);
    terms[4] = s[8] / (3.0f * (float)HW);
  }
}

// g_terms[5] = upstream gradients of (head, eyes, face, nonhead, bg)  ->  g_img_face, g_img_eyes, g_img [B,3,HW], g_bg_img [1,3,HW]
__global__ void __launch_bounds__(256) data_loss_bwd_kernel(const LossArgs a, const float* __restrict__ sums, const float* __restrict__ g_terms,
                                                            float* __restrict__ g_img_face, float* __restrict__ g_img_eyes,
                                                            float* __restrict__ g_img, float* __restrict__ g_bg_img) {
  const long long total = (long long)a.B * a.HW;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int b = (int)(idx / a.HW), p = (int)(idx - (long long)b * a.HW);
  float m_head, m_face, m_eyes, m_nh;
  pixel_masks(a, (size_t)idx, m_head, m_face, m_eyes, m_nh);
  const float k_head = g_terms[0] / (3.0f * sums[1]), k_eyes = g_terms[1] / (3.0f * sums[3]), k_face = g_terms[2] / (3.0f * sums[5]);
  const float k_nh = g_terms[3] / (3.0f * sums[7]), k_bg = g_terms[4] / (3.0f * (float)a.HW);
  auto dd = [&](float d) { return a.use_l1 ? (d > 0.0f ? 1.0f : (d < 0.0f ? -1.0f : 0.0f)) : 2.0f * d; };   // torch: sign(0) = 0
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const size_t o = ((size_t)b * 3 + c) * a.HW + p;
    const float g = a.gt[o];
    const float v = a.img[o];
    g_img[o] = (m_head > 0.0f ? k_head * dd(v - g) : 0.0f) + (m_nh > 0.0f ? k_nh * 2.0f * (v - a.bg_value) : 0.0f);
    g_img_eyes[o] = m_eyes > 0.0f ? k_eyes * dd(a.img_eyes[o] - g) : 0.0f;
    g_img_face[o] = m_face > 0.0f ? k_face * dd(a.img_face[o] - g) : 0.0f;
    if (b == 0) g_bg_img[(size_t)c * a.HW + p] = k_bg * 2.0f * (a.bg_img[(size_t)c * a.HW + p] - a.bg_value);
  }
}

static int fill_args(LossArgs& a, const float* img_face, const float* img_eyes, const float* img, const float* bg_img, const float* gt,
                     const float* face_mask, const float* full_eye, const float* left_eye, const float* right_eye, int B, int HW, int use_l1,
                     float bg_value) {
  a.img_face = img_face; a.img_eyes = img_eyes; a.img = img; a.bg_img = bg_img; a.gt = gt;
  a.face_mask = face_mask; a.full_eye = full_eye; a.left_eye = left_eye; a.right_eye = right_eye;
  a.B = B; a.HW = HW; a.use_l1 = use_l1; a.bg_value = bg_value;
  return GNRF_OK;
}

constexpr int kLossBlocks = 296;   // 2 x 148 SMs

}  // namespace gnrf

using namespace gnrf;

extern "C" size_t gnrf_data_loss_workspace_floats(void) { return (size_t)kLossBlocks * kLossSums; }

extern "C" int gnrf_data_loss_fwd(const float* img_face, const float* img_eyes, const float* img, const float* bg_img, const float* gt,
                                  const float* face_mask, const float* full_eye, const float* left_eye, const float* right_eye, int B, int HW,
                                  int use_l1, float bg_value, float* terms, float* sums, float* workspace, gnrf_stream_t stream) {
  GNRF_CHECK_ARG(img_face && img_eyes && img && bg_img && gt && face_mask && full_eye && left_eye && right_eye && terms && sums && workspace);
  GNRF_CHECK_ARG(B > 0 && HW > 0);
  LossArgs a;
  fill_args(a, img_face, img_eyes, img, bg_img, gt, face_mask, full_eye, left_eye, right_eye, B, HW, use_l1, bg_value);
  data_loss_partial_kernel<<<kLossBlocks, 256, 0, as_stream(stream)>>>(a, workspace);
  data_loss_final_kernel<<<1, 32, 0, as_stream(stream)>>>(workspace, kLossBlocks, HW, sums, terms);
  GNRF_LAUNCH_CHECK();
  count_launches(2);
  return GNRF_OK;
}

extern "C" int gnrf_data_loss_bwd(const float* img_face, const float* img_eyes, const float* img, const float* bg_img, const float* gt,
                                  const float* face_mask, const float* full_eye, const float* left_eye, const float* right_eye, int B, int HW,
                                  int use_l1, float bg_value, const float* sums, const float* g_terms, float* g_img_face, float* g_img_eyes,
                                  float* g_img, float* g_bg_img, gnrf_stream_t stream) {
  GNRF_CHECK_ARG(img_face && img_eyes && img && bg_img && gt && face_mask && full_eye && left_eye && right_eye && sums && g_terms);
  GNRF_CHECK_ARG(g_img_face && g_img_eyes && g_img && g_bg_img && B > 0 && HW > 0);
  LossArgs a;
  fill_args(a, img_face, img_eyes, img, bg_img, gt, face_mask, full_eye, left_eye, right_eye, B, HW, use_l1, bg_value);
  const long long total = (long long)B * HW;
  data_loss_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(a, sums, g_terms, g_img_face, g_img_eyes, g_img, g_bg_img);
  GNRF_LAUNCH_CHECK();
  count_launches(1);
  return GNRF_OK;
}
