// Backward of the 2-D neural renderer (training path).
// Differentiates NeuralRenderer.forward (models/neural_renderer.py:98-113) and PixelShuffleUpsample.forward
// (models/pixel_shuffle_upsample.py:19-42) as evaluated by gnrf_nr_train_fwd:
//   rgb_0 = toRGB_0(x);  per level i:  t1 = LReLU(W1 net_i), sh = shuffle(LReLU(W2 t1) + repeat(net_i,4)),
//   net_{i+1} = LReLU(Blur(W3 sh + b3)),  rgb_{i+1} = Blur(Up2(rgb_i)) + toRGB_{i+1}(net_{i+1});  img = sigmoid(rgb_last).
// Every GEMM runs on tcgen05: input gradients with conv_tc (W^T packed here), weight gradients with wgrad_tc; the fixed linear
// stages (blur, bilinear x2, pixel shuffle) use their adjoint kernels from train_ops.cu.
#include "common.cuh"
#include "conv_tc.cuh"
#include "nr_plan.cuh"
#include "train_ops.cuh"
#include "wgrad_tc.cuh"

namespace gnrf {

struct NrBwdPlan {
  size_t packT, wg, gbuf[7], grgb[8], grgb_tmp, total;   // byte offsets
  size_t wg_bytes, gbuf_floats;
};

static inline int nr_width(int C, int i, int min_feat) { return (C >> i) > min_feat ? (C >> i) : min_feat; }

static NrBwdPlan nr_bwd_plan(int N, int C, int S, int n_blocks, int min_feat) {
  NrBwdPlan p;
  auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
  size_t packT = 0, wg = 0, gmax = 0;
  auto wgb = [&](int Ndy, int Kx, int HW) {
    size_t a = tc::wgrad_plan(Ndy, Kx, N, HW, true).partial_bytes;
    if (a > wg) wg = a;
  };
  for (int i = 0; i < n_blocks; ++i) {
    int ci = nr_width(C, i, min_feat), co = nr_width(C, i + 1, min_feat);
    int s = S << i;
    packT += tc::conv_layer_plan(ci, 2 * ci).total_bytes + tc::conv_layer_plan(2 * ci, 4 * ci).total_bytes +
             tc::conv_layer_plan(ci, co).total_bytes;
    wgb(2 * ci, ci, s * s); wgb(4 * ci, 2 * ci, s * s); wgb(co, ci, 4 * s * s); wgb(3, co, 4 * s * s);
    gmax = max(gmax, (size_t)N * ci * 4 * s * s);
  }
  wgb(3, C, S * S);
  for (int j = 0; j <= n_blocks; ++j) packT += tc::conv_layer_plan(j == 0 ? C : nr_width(C, j, min_feat), 3).total_bytes;
  size_t off = 0;
  p.packT = off; off += al(packT);
  p.wg = off; off += al(wg); p.wg_bytes = wg;
  p.gbuf_floats = gmax;
  for (int k = 0; k < 7; ++k) { p.gbuf[k] = off; off += al(gmax * sizeof(float)); }
  for (int j = 0; j <= n_blocks; ++j) {
    size_t s = (size_t)S << j;
    p.grgb[j] = off; off += al((size_t)N * 3 * s * s * sizeof(float));
  }
  size_t P = (size_t)S << n_blocks;
  p.grgb_tmp = off; off += al((size_t)N * 3 * P * P * sizeof(float));
  p.total = off;
  return p;
}

}  // namespace gnrf

using namespace gnrf;

extern "C" size_t gnrf_nr_train_bwd_workspace_bytes(int N, int C, int S, int n_blocks, int min_feat) {
  if (N <= 0 || C <= 0 || S <= 0 || n_blocks < 1 || n_blocks > 6) return 0;
  return nr_bwd_plan(N, C, S, n_blocks, min_feat).total;
}

extern "C" int gnrf_nr_train_bwd(const float* const* params, const float* featmap, const void* saved, const float* img, const float* g_img,
                                 int N, int C, int S, int n_blocks, int min_feat, float* g_featmap, float* const* g_params,
                                 void* workspace, size_t workspace_bytes, gnrf_stream_t stream) {
  GNRF_CHECK_ARG(params && featmap && saved && img && g_img && g_featmap && g_params && workspace);
  GNRF_CHECK_ARG(N > 0 && C > 0 && S >= 2 && n_blocks >= 1 && n_blocks <= 6 && S % 4 == 0);
  for (int i = 0; i < 8 * n_blocks + 2; ++i) GNRF_CHECK_ARG(params[i] != nullptr && g_params[i] != nullptr);
  const NrTrainPlan sp = nr_train_plan(N, C, S, n_blocks, min_feat);
  const NrBwdPlan bp = nr_bwd_plan(N, C, S, n_blocks, min_feat);
  if (workspace_bytes < bp.total)
    return fail(GNRF_ERR_ARG, "gnrf_nr_train_bwd: workspace %zu < required %zu bytes", workspace_bytes, bp.total);
  cudaStream_t st = as_stream(stream);
  const char* sv = static_cast<const char*>(saved);
  char* ws = static_cast<char*>(workspace);
  auto F = [&](size_t off) { return reinterpret_cast<float*>(ws + off); };
  // parameter indexing (gnrf.h): psu i: [4i..4i+3]; to_rgb j: [4nb + 2j, +1]; feat i: [4nb + 2(nb+1) + 2i, +1]
  const int nb = n_blocks;
  auto i_psu = [&](int i, int l) { return 4 * i + 2 * l; };
  auto i_rgb = [&](int j) { return 4 * nb + 2 * j; };
  auto i_feat = [&](int i) { return 4 * nb + 2 * (nb + 1) + 2 * i; };
  void* wg = ws + bp.wg;
  unsigned char* packT = reinterpret_cast<unsigned char*>(ws + bp.packT);
  size_t pk_off = 0;
  int rc;
  // input gradient of a conv: out = W^T g  (+ mask / add), W is the forward weight [Nf][Kf]  ->  GEMM with N = Kf, K = Nf
  auto conv_dx = [&](const float* Wf, int Nf, int Kf, const float* g, float* out, int HW, const float* mask, float slope, const float* add) {
    tc::ConvLayerPlan pl = tc::conv_layer_plan(Kf, Nf);
    unsigned char* dst = packT + pk_off;
    pk_off += pl.total_bytes;
    int r = tc::conv_tc_pack_strided(pl, Wf, 1, Kf, nullptr, dst, st);
    if (r != GNRF_OK) return r;
    tc::ConvExtras ex = {};
    ex.mask = mask; ex.mask_slope = slope; ex.add = add;
    return tc::conv_tc_launch_ex(pl, dst, g, out, nullptr, 1, N, HW, HW, tc::CONV_EPI_LINEAR, ex, st);
  };
  auto wgrad = [&](const float* g, int Nf, const float* X, int Kf, int HW, int pidx) {
    return tc::wgrad_tc_launch(g, (long long)Nf * HW, X, (long long)Kf * HW, Nf, Kf, N, HW, g_params[pidx], g_params[pidx + 1], 1, 0, wg,
                               bp.wg_bytes, st);
  };

  const size_t Pimg = (size_t)S << nb;
  // g_rgb[nb] = sigmoid'(img) * g_img
  launch_sigmoid_bwd(g_img, img, (long long)N * 3 * Pimg * Pimg, F(bp.grgb[nb]), st);
  float* g_net = nullptr;   // gradient w.r.t. the output of level i (net_{i+1}) coming from level i+1; null at the last level
  for (int i = nb - 1; i >= 0; --i) {
    const int ci = nr_width(C, i, min_feat), co = nr_width(C, i + 1, min_feat);
    const int s = S << i, s2 = 2 * s;
    const int HW = s * s, HW2 = s2 * s2;
    const float* net_in = i == 0 ? featmap : reinterpret_cast<const float*>(sv + sp.net[i - 1]);
    const float* t1 = reinterpret_cast<const float*>(sv + sp.t1[i]);
    const float* sh = reinterpret_cast<const float*>(sv + sp.sh[i]);
    const float* net_out = reinterpret_cast<const float*>(sv + sp.net[i]);
    float* g_rgb = F(bp.grgb[i + 1]);
    // to-RGB head i+1: weight grads, then g_net_total = W_rgb^T g_rgb (+ g_net from the level above)
    if ((rc = wgrad(g_rgb, 3, net_out, co, HW2, i_rgb(i + 1))) != GNRF_OK) return rc;
    float* g_net_tot = F(bp.gbuf[0]);      // g_net (if any) lives in gbuf[1]
    if ((rc = conv_dx(params[i_rgb(i + 1)], 3, co, g_rgb, g_net_tot, HW2, nullptr, 1.0f, g_net)) != GNRF_OK) return rc;
    // rgb_{i+1} = Blur(Up2(rgb_i)) + ...  ->  g_rgb_i = Up2^T Blur^T g_rgb_{i+1}
    launch_blur_adj(g_rgb, nullptr, 1.0f, N * 3, s2, s2, F(bp.grgb_tmp), st);
    launch_up2_adj(F(bp.grgb_tmp), N * 3, s, s, F(bp.grgb[i]), st);
    // net_{i+1} = LReLU(Blur(pre)):  g_pre = Blur^T (g_net_tot * slope(net_{i+1}))
    float* g_blpre = F(bp.gbuf[2]);
    launch_blur_adj(g_net_tot, net_out, 0.2f, N * co, s2, s2, g_blpre, st);
    if ((rc = wgrad(g_blpre, co, sh, ci, HW2, i_feat(i))) != GNRF_OK) return rc;
    float* g_sh = F(bp.gbuf[3]);
    if ((rc = conv_dx(params[i_feat(i)], co, ci, g_blpre, g_sh, HW2, nullptr, 1.0f, nullptr)) != GNRF_OK) return rc;
    float* g_pre2 = F(bp.gbuf[4]);
    float* g_res = F(bp.gbuf[5]);
    launch_psu_bwd(g_sh, sh, net_in, N, ci, s, s, g_pre2, g_res, st);
    if ((rc = wgrad(g_pre2, 4 * ci, t1, 2 * ci, HW, i_psu(i, 1))) != GNRF_OK) return rc;
    float* g_pre1 = F(bp.gbuf[6]);
    if ((rc = conv_dx(params[i_psu(i, 1)], 4 * ci, 2 * ci, g_pre2, g_pre1, HW, t1, 0.2f, nullptr)) != GNRF_OK) return rc;
    if ((rc = wgrad(g_pre1, 2 * ci, net_in, ci, HW, i_psu(i, 0))) != GNRF_OK) return rc;
    float* g_net_in = F(bp.gbuf[1]);       // the old g_net (consumed above) lived here
    if ((rc = conv_dx(params[i_psu(i, 0)], 2 * ci, ci, g_pre1, g_net_in, HW, nullptr, 1.0f, g_res)) != GNRF_OK) return rc;
    g_net = g_net_in;
    GNRF_LAUNCH_CHECK();
  }
  // head 0 on the input feature map
  if ((rc = wgrad(F(bp.grgb[0]), 3, featmap, C, S * S, i_rgb(0))) != GNRF_OK) return rc;
  if ((rc = conv_dx(params[i_rgb(0)], 3, C, F(bp.grgb[0]), g_featmap, S * S, nullptr, 1.0f, g_net)) != GNRF_OK) return rc;
  GNRF_LAUNCH_CHECK();
  return GNRF_OK;
}
