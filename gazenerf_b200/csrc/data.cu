// Dataset sample -> device tensors: the per-item transforms of ETH-XGaze __getitem__
// (datasets/eth_xgaze.py:308-360) and GazeNeRFTrainer.prepare_data (trainer/gazenerf_trainer.py:250-337) on the raw HDF5 arrays
// (schema dataset_pre_processing.py:260-380), so that only the u8 / f64 records cross PCIe (0.98 MB per sample instead of the 6 MB of
// fp32 tensors the reference uploads) and the CPU does no per-pixel work.
//
//   face_patch u8 [B,H,W,3] BGR -> img f32 [B,3,H,W] RGB / 255         (image[:, :, [2,1,0]]; ToPILImage -> ToTensor, :12,:331-332)
//   head_mask  u8 [B,H,W]       -> cv2.erode(3x3 ones, iterations=2)   (:338-339) -> f32 [B,1,H,W]
//   left/right eye mask u8      -> f32 [B,1,H,W]
//   latent_codes f64: code = row 0 of the subject with [279:] taken from the sample's row (:346-347) -> iden100 | expr79 | text100 | illu27 f32
//   pitchyaw_head f64 [B,2] -> gaze f32; c2w_Rmat / c2w_Tvec f64 -> f32 [B,3,3] / [B,3,1]
//   inmat f64 [B,3,3]: rows 0,1 *= featmap/img size, closed-form inverse in f64, cast to f32 (trainer :317-336)
// HBM-bound streaming kernels: 1 byte in / 4 bytes out per plane element; the 5x5 erosion window is served by L1/L2.
#include "common.cuh"

namespace gnrf {

constexpr int kPxPerThread = 4;

// cv2.erode(mask, ones(3,3), iterations = it) == (2 it + 1)^2 minimum; out-of-image taps are ignored (cv2's default border value for
// erosion is +inf).
__global__ void sample_images_kernel(const uint8_t* __restrict__ face, const uint8_t* __restrict__ head, const uint8_t* __restrict__ left,
                                     const uint8_t* __restrict__ right, int B, int H, int W, int erode_r, float* __restrict__ img,
                                     float* __restrict__ head_f, float* __restrict__ left_f, float* __restrict__ right_f) {
  const int wq = W / kPxPerThread;
  const long long total = (long long)B * H * wq;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int xq = (int)(t % wq);
    const int y = (int)((t / wq) % H);
    const int b = (int)(t / ((long long)wq * H));
    const int x0 = xq * kPxPerThread;
    const size_t pix = ((size_t)b * H + y) * W + x0;
    // ---- image: 4 pixels = 12 bytes BGR, 4-byte aligned because W % 4 == 0
    const uint32_t* p = reinterpret_cast<const uint32_t*>(face + pix * 3);
    const uint32_t w0 = __ldg(p), w1 = __ldg(p + 1), w2 = __ldg(p + 2);
    const uint8_t by[12] = {(uint8_t)w0, (uint8_t)(w0 >> 8), (uint8_t)(w0 >> 16), (uint8_t)(w0 >> 24), (uint8_t)w1, (uint8_t)(w1 >> 8),
                            (uint8_t)(w1 >> 16), (uint8_t)(w1 >> 24), (uint8_t)w2, (uint8_t)(w2 >> 8), (uint8_t)(w2 >> 16), (uint8_t)(w2 >> 24)};
    const size_t plane = (size_t)H * W;
    float* o = img + (size_t)b * 3 * plane + (size_t)y * W + x0;
#pragma unroll
    for (int c = 0; c < 3; ++c) {   // output channel c (R,G,B) = input byte 2 - c of every BGR triple
      float4 v;
      v.x = __fdiv_rn((float)by[0 + 2 - c], 255.0f);
      v.y = __fdiv_rn((float)by[3 + 2 - c], 255.0f);
      v.z = __fdiv_rn((float)by[6 + 2 - c], 255.0f);
      v.w = __fdiv_rn((float)by[9 + 2 - c], 255.0f);
      *reinterpret_cast<float4*>(o + c * plane) = v;
    }
    // ---- eye masks: plain u8 -> f32
    {
      const uint32_t l = __ldg(reinterpret_cast<const uint32_t*>(left + pix)), r = __ldg(reinterpret_cast<const uint32_t*>(right + pix));
      *reinterpret_cast<float4*>(left_f + pix) = make_float4((float)(l & 255u), (float)((l >> 8) & 255u), (float)((l >> 16) & 255u), (float)(l >> 24));
      *reinterpret_cast<float4*>(right_f + pix) = make_float4((float)(r & 255u), (float)((r >> 8) & 255u), (float)((r >> 16) & 255u), (float)(r >> 24));
    }
    // ---- head mask: (2r+1)^2 erosion
    {
      unsigned int m[kPxPerThread] = {255u, 255u, 255u, 255u};
      const uint8_t* hb = head + (size_t)b * plane;
      for (int dy = -erode_r; dy <= erode_r; ++dy) {
        const int yy = y + dy;
        if (yy < 0 || yy >= H) continue;
        const uint8_t* row = hb + (size_t)yy * W;
        for (int xx = x0 - erode_r; xx < x0 + kPxPerThread + erode_r; ++xx) {
          if (xx < 0 || xx >= W) continue;
          const unsigned int v = __ldg(row + xx);
#pragma unroll
          for (int j = 0; j < kPxPerThread; ++j)
            if (xx >= x0 + j - erode_r && xx <= x0 + j + erode_r) m[j] = min(m[j], v);
        }
      }
      *reinterpret_cast<float4*>(head_f + pix) = make_float4((float)m[0], (float)m[1], (float)m[2], (float)m[3]);
    }
  }
}

// one thread per (sample, output element)
__global__ void sample_meta_kernel(const double* __restrict__ code_row0, const double* __restrict__ code_rows, const double* __restrict__ pitchyaw,
                                   const double* __restrict__ c2w_R, const double* __restrict__ c2w_T, const double* __restrict__ inmat, int B,
                                   double k_scale, float* __restrict__ iden, float* __restrict__ expr, float* __restrict__ text,
                                   float* __restrict__ illu, float* __restrict__ gaze, float* __restrict__ R, float* __restrict__ T,
                                   float* __restrict__ Kinv) {
  const int b = blockIdx.x;
  for (int i = threadIdx.x; i < 306; i += blockDim.x) {
    const double v = (i < 279) ? code_row0[i] : code_rows[(size_t)b * 306 + i];   // eth_xgaze.py:346-347
    const float f = (float)v;
    if (i < 100) iden[b * 100 + i] = f;
    else if (i < 179) expr[b * 79 + (i - 100)] = f;
    else if (i < 279) text[b * 100 + (i - 179)] = f;
    else illu[b * 27 + (i - 279)] = f;
  }
  if (threadIdx.x < 2) gaze[b * 2 + threadIdx.x] = (float)pitchyaw[b * 2 + threadIdx.x];
  if (threadIdx.x < 9) R[b * 9 + threadIdx.x] = (float)c2w_R[b * 9 + threadIdx.x];
  if (threadIdx.x < 3) T[b * 3 + threadIdx.x] = (float)c2w_T[b * 3 + threadIdx.x];
  if (threadIdx.x == 0) {
    // temp_inmat[:, :2, :] *= featmap / img ; closed-form inverse (trainer/gazenerf_trainer.py:317-325), all in f64
    const double* k = inmat + (size_t)b * 9;
    const double fx = k[0] * k_scale, fy = k[4] * k_scale, cx = k[2] * k_scale, cy = k[5] * k_scale;
    float* o = Kinv + b * 9;
    o[0] = (float)(1.0 / fx); o[1] = 0.0f; o[2] = (float)(-(cx / fx));
    o[3] = 0.0f; o[4] = (float)(1.0 / fy); o[5] = (float)(-(cy / fy));
    o[6] = 0.0f; o[7] = 0.0f; o[8] = 1.0f;
  }
}

}  // namespace gnrf

using namespace gnrf;

extern "C" int gnrf_sample_images_to_device(const uint8_t* face_patch_bgr, const uint8_t* head_mask, const uint8_t* left_eye_mask,
                                            const uint8_t* right_eye_mask, int B, int H, int W, int erode_iterations, float* img,
                                            float* head_mask_f, float* left_eye_mask_f, float* right_eye_mask_f, gnrf_stream_t stream) {
  GNRF_CHECK_ARG(face_patch_bgr && head_mask && left_eye_mask && right_eye_mask && img && head_mask_f && left_eye_mask_f && right_eye_mask_f);
  GNRF_CHECK_ARG(B > 0 && H > 0 && W > 0 && W % 4 == 0 && erode_iterations >= 0 && erode_iterations <= 8);
  GNRF_CHECK_ARG((reinterpret_cast<uintptr_t>(face_patch_bgr) & 3) == 0 && (reinterpret_cast<uintptr_t>(left_eye_mask) & 3) == 0 &&
                 (reinterpret_cast<uintptr_t>(right_eye_mask) & 3) == 0);
  const long long total = (long long)B * H * (W / kPxPerThread);
  const int threads = 256;
  long long blocks = (total + threads - 1) / threads;
  if (blocks > 148 * 16) blocks = 148 * 16;
  sample_images_kernel<<<(int)blocks, threads, 0, as_stream(stream)>>>(face_patch_bgr, head_mask, left_eye_mask, right_eye_mask, B, H, W,
                                                                       erode_iterations, img, head_mask_f, left_eye_mask_f, right_eye_mask_f);
  GNRF_LAUNCH_CHECK();
  count_launches(1);
  return GNRF_OK;
}

extern "C" int gnrf_sample_meta_to_device(const double* code_row0, const double* code_rows, const double* pitchyaw, const double* c2w_Rmat,
                                          const double* c2w_Tvec, const double* inmat, int B, int featmap_size, int img_size, float* iden,
                                          float* expr, float* text, float* illu, float* gaze, float* rmats, float* tvecs, float* inv_inmats,
                                          gnrf_stream_t stream) {
  GNRF_CHECK_ARG(code_row0 && code_rows && pitchyaw && c2w_Rmat && c2w_Tvec && inmat && iden && expr && text && illu && gaze && rmats && tvecs &&
                 inv_inmats);
  GNRF_CHECK_ARG(B > 0 && featmap_size > 0 && img_size > 0);
  sample_meta_kernel<<<B, 128, 0, as_stream(stream)>>>(code_row0, code_rows, pitchyaw, c2w_Rmat, c2w_Tvec, inmat, B,
                                                       (double)featmap_size / (double)img_size, iden, expr, text, illu, gaze, rmats, tvecs,
                                                       inv_inmats);
  GNRF_LAUNCH_CHECK();
  count_launches(1);
  return GNRF_OK;
}
