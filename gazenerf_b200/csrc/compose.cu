// Feature-map compose: background blend, gaze rotation of channel triplets, max-merge (HBM-bound, one pass).
// Reference: models/gaze_nerf.py:175-203; rotation_matrix_2d / rotate, utils/model_utils.py:11-46.
#include "common.cuh"

namespace gnrf {

// One thread per (b, triplet k, pixel p).  Reads 6 feature values + 2 alphas + 3 bg values, writes 9.
__global__ void compose_kernel(const float* __restrict__ feat_face, const float* __restrict__ a_face,
                               const float* __restrict__ feat_eyes, const float* __restrict__ a_eyes,
                               const float* __restrict__ bg, const float* __restrict__ gaze, int B, int C, int P,
                               float* __restrict__ out) {
  const int n_trip = C / 3;
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)B * n_trip * P;
  if (idx >= total) return;
  int p = (int)(idx % P);
  int k = (int)((idx / P) % n_trip);
  int b = (int)(idx / ((long long)P * n_trip));

  // R = Ry(g1) * Rx(g0)  (utils/model_utils.py:11-26)
  float s0, c0, s1, c1;
  sincosf(gaze[b * 2 + 0], &s0, &c0);
  sincosf(gaze[b * 2 + 1], &s1, &c1);
  // Rx = [[1,0,0],[0,c0,-s0],[0,s0,c0]], Ry = [[c1,0,s1],[0,1,0],[-s1,0,c1]]
  float R[3][3];
  R[0][0] = c1;  R[0][1] = s1 * s0;  R[0][2] = s1 * c0;
  R[1][0] = 0.f; R[1][1] = c0;       R[1][2] = -s0;
  R[2][0] = -s1; R[2][1] = c1 * s0;  R[2][2] = c1 * c0;

  float af = a_face[(size_t)b * P + p], ae = a_eyes[(size_t)b * P + p];
  float mf[3], me[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    size_t ch = (size_t)(3 * k + i);
    float bgv = bg[ch * P + p];
    mf[i] = fmaf(af, bgv, feat_face[((size_t)b * C + ch) * P + p]);  // fg + bg_alpha * bg (models/gaze_nerf.py:178-179)
    me[i] = fmaf(ae, bgv, feat_eyes[((size_t)b * C + ch) * P + p]);
  }
  const size_t plane = (size_t)B * C * P;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    // row-vector times R: out_j = sum_i v_i R[i][j]  (utils/model_utils.py:41-43)
    float ep = fmaf(me[2], R[2][j], fmaf(me[1], R[1][j], me[0] * R[0][j]));
    size_t o = ((size_t)b * C + (3 * k + j)) * P + p;
    out[o] = mf[j];
    out[plane + o] = ep;
    out[2 * plane + o] = fmaxf(mf[j], ep);  // models/gaze_nerf.py:203
  }
}

}  // namespace gnrf

using namespace gnrf;

extern "C" int gnrf_compose_fwd(const float* feat_face, const float* a_face, const float* feat_eyes, const float* a_eyes,
                                const float* bg, const float* gaze, int B, int C, int P, float* out, gnrf_stream_t stream) {
  GNRF_CHECK_ARG(feat_face && a_face && feat_eyes && a_eyes && bg && gaze && out);
  GNRF_CHECK_ARG(B > 0 && C > 0 && P > 0);
  if (C % 3 != 0) return fail(GNRF_ERR_ARG, "gnrf_compose_fwd: featmap_nc=%d must be a multiple of 3", C);
  long long total = (long long)B * (C / 3) * P;
  compose_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(feat_face, a_face, feat_eyes, a_eyes, bg, gaze,
                                                                               B, C, P, out);
  GNRF_LAUNCH_CHECK();
  count_launches(1);
  return GNRF_OK;
}
