/* gnrf.h -- C ABI of the B200-native GazeNeRF render hot path (libgnrf.so).
 *
 * The reference (AlessandroRuzzi/GazeNeRF) is 100 % Python/PyTorch and has no native boundary; the boundary
 * it DOES have is the module call  GazeNeRFNet.forward(mode, batch_xy, batch_uv, bg_code, shape_code,
 * appea_code, gaze_code, batch_Rmats, batch_Tvecs, batch_inv_inmats)  (models/gaze_nerf.py:322-351).  Each
 * entry point below replaces one stage of that call; the reference file:line it replaces is cited per
 * function.  gazenerf_b200/net.py binds these with ctypes and re-creates the module API on top.
 *
 * Conventions (all functions)
 *   - plain pointers + ints only; every pointer is a DEVICE pointer on the current CUDA device unless the
 *     parameter is documented "host"; float = IEEE fp32; tensors are dense, row-major in the stated shape.
 *   - no allocation, no ownership transfer: the caller owns every input, output and workspace buffer.
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*); calls are asynchronous and re-entrant: the library keeps no
 *     device-global tables and reads no environment variables; the only process state is the launch counter, the thread-local
 *     error string and a mutex-guarded per-device "shared-memory opt-in done" bitmask (any device, any thread may call first).
 *   - return 0 on success, a GNRF_ERR_* code otherwise; gnrf_last_error() returns a thread-local message.
 *     No exception or abort ever crosses the boundary (the reference's train loop swallows exceptions per
 *     batch, trainer/gazenerf_trainer.py:576-582 -- the Python wrapper raises RuntimeError from the code).
 *
 * Not covered (by design): gradients with the view-direction input (include_vd=True is inference-only here, see
 * gnrf_mlp_tc_fwd_vd; every reference entry point constructs the network with include_vd=False) and a backward for the
 * hierarchical (FineSample) pass (the reference's own hier wiring is dead code, models/gaze_nerf.py:282-318).
 */
#ifndef GNRF_H_
#define GNRF_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GNRF_ABI_VERSION 1

enum {
  GNRF_OK = 0,
  GNRF_ERR_ARG = 1,         /* bad shape / null pointer / unsupported size */
  GNRF_ERR_CUDA = 2,        /* a CUDA runtime call or launch failed */
  GNRF_ERR_UNSUPPORTED = 3  /* configuration outside what the sm_100a kernels are specialised for */
};

/* Fixed network geometry of the path (models/gaze_nerf.py:22-62, configs/gazenerf_options.py:11-26). */
#define GNRF_PE_DIMS 63          /* 3 + 6*10 positional-encoding channels */
#define GNRF_SHAPE_EXT_DIMS 181  /* iden 100 + expr 79 + gaze 2 */
#define GNRF_APPEA_DIMS 127      /* text 100 + illu 27 */
#define GNRF_MLP_NPARAMS 24      /* 12 layers x (weight, bias), state_dict order */

typedef void* gnrf_stream_t; /* cudaStream_t */

int gnrf_abi_version(void);
const char* gnrf_last_error(void);
/* 0 if the current device is sm_100 (B200) and the kernels can run, else GNRF_ERR_UNSUPPORTED. */
int gnrf_device_check(void);
/* Number of libgnrf kernels launched by this process so far (monotonic; used for bench.py's "gpu_launches"). */
unsigned long long gnrf_launch_count(void);

/* ---------------------------------------------------------------------------------------------------------
 * Ray generation.  Replaces GenSamplePoints.forward, utils/model_utils.py:364-372.
 *   xy [B,2,N_r] (row 0 = x, row 1 = y), rmats [B,3,3] cam-to-world, inv_inmats [B,3,3]
 *   -> ray_dl [B,N_r,4] = (d_x, d_y, d_z, l) with d = normalize(R K^-1 (x,y,1)), l = -1/d_z.
 * (ray origin o = Tvec is passed separately to the consumers.) */
int gnrf_ray_setup(const float* xy, const float* rmats, const float* inv_inmats, int B, int N_r, float* ray_dl,
                   gnrf_stream_t stream);

/* Coarse depth edges.  Replaces GenSamplePoints._calc_sample_points + the jitter of
 * _calc_sample_points_by_zvals, utils/model_utils.py:332-357, 302-307.
 *   tvecs [B,3]; t_vals [N_s+1] = linspace(0,1,N_s+1) supplied by the host (keeps torch's rounding);
 *   jitter_u [B,N_r,N_s+1] uniform draws (train mode) or NULL (eval)
 *   -> z_edges [B,N_r,N_s+1];  sample k has depth z_edges[k] and extent (z_edges[k+1]-z_edges[k]) * l. */
int gnrf_coarse_depths(const float* tvecs, const float* t_vals, const float* jitter_u, int B, int N_r, int N_s,
                       float world_z1, float world_z2, float* z_edges, gnrf_stream_t stream);

/* Hierarchical (inverse-CDF) fine sampling.  Replaces FineSample.forward, utils/model_utils.py:404-477.
 *   weights [B,N_r,N_c] coarse compositing weights; z_edges_coarse [B,N_r,N_c+1];
 *   u: [N_f1] shared by all rays when u_per_ray == 0, else [B*N_r, N_f1]   (N_f1 = num_sample_fine + 1)
 *   -> inds (int64 [B*N_r, N_f1], may be NULL) = searchsorted(cdf, u, right=True), bit-exact integer work;
 *      z_edges_fine [B,N_r,N_c+N_f1] = sort(cat(coarse sample depths, fine depths)). */
int gnrf_fine_depths(const float* weights, const float* z_edges_coarse, const float* u, int u_per_ray, int B, int N_r,
                     int N_c, int N_f1, int64_t* inds, float* z_edges_fine, gnrf_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * Radiance MLP, exact-fp32 CUDA-core path (one fused kernel per branch: ray points -> positional encoding
 * -> 12 dense layers).  Replaces Embedder.forward (utils/model_utils.py:272-280), the code broadcast/concat
 * (models/gaze_nerf.py:248-262,136-143) and MLPforNeRF.forward (models/mlp_nerf.py:95-119), literally (no
 * algebraic folds).  params: HOST array of 24 device pointers, state_dict order:
 *   FeaExt_module_0..7 (w,b), density_module (w,b), RGB_layer_0 (w,b), RGB_layer_1 (w,b), RGB_layer_2 (w,b).
 *   -> feat_pts [B,N_r,N_s,n_feat], sigma_pts [B,N_r,N_s] (post-ReLU). */
int gnrf_mlp_simt_fwd(const float* const* params, const float* ray_dl, const float* tvecs, const float* z_edges,
                      const float* shape_ext, const float* appea, int B, int N_r, int N_s, int hidden, int n_feat,
                      float* feat_pts, float* sigma_pts, gnrf_stream_t stream);

/* Alpha compositing.  Replaces CalcRayColor.forward, utils/model_utils.py:493-534.
 *   -> feat_ray [B,n_feat,N_r], bg_alpha [B,N_r], depth [B,N_r] (nullable), weights [B,N_r,N_s] (nullable). */
int gnrf_composite_fwd(const float* feat_pts, const float* sigma_pts, const float* z_edges, const float* ray_dl, int B,
                       int N_r, int N_s, int n_feat, float* feat_ray, float* bg_alpha, float* depth, float* weights,
                       gnrf_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * Radiance MLP, tcgen05 tensor-core path (sm_100a): ONE fused kernel for both branches doing
 * ray points -> PE -> dense stack (bf16x3 split-precision UMMA, fp32 accumulate in TMEM, weights streamed by
 * TMA bulk copies) -> ReLU density -> per-ray alpha composite.  Same reference lines as the two entries above.
 * Exact algebraic folds used: per-face code columns folded into bias vectors, density row
 * appended to the RGB head, RGB_layer_0*RGB_layer_1 pre-multiplied, RGB_layer_2 applied after compositing.
 *
 * gnrf_mlp_tc_pack: params (host array of 24 device ptrs) -> packed (device, gnrf_mlp_tc_packed_bytes()).
 * gnrf_mlp_tc_fold: per-face bias block [B, gnrf_mlp_tc_bias_floats()] from the packed fold matrices + codes.
 * gnrf_mlp_tc_fwd : n_branch in {1,2}; packed[i], bias[i] per branch (host arrays of device pointers)
 *   -> feat_ray[i] [B,n_feat,N_r], bg_alpha[i] [B,N_r], weights[i] [B,N_r,N_s] (array or entries may be NULL).
 * N_s must divide 128 (1 tile = 128 consecutive samples = 128/N_s whole rays). */
size_t gnrf_mlp_tc_packed_bytes(void);
size_t gnrf_mlp_tc_bias_floats(void);
int gnrf_mlp_tc_pack(const float* const* params, void* packed, gnrf_stream_t stream);
int gnrf_mlp_tc_fold(const void* packed, const float* shape_ext, const float* appea, int B, float* bias,
                     gnrf_stream_t stream);
size_t gnrf_mlp_tc_workspace_bytes(int n_branch, int B, int N_r);
int gnrf_mlp_tc_fwd(int n_branch, const void* const* packed, const float* const* bias, const float* ray_dl,
                    const float* tvecs, const float* z_edges, int B, int N_r, int N_s, float* const* feat_ray,
                    float* const* bg_alpha, float* const* weights, void* workspace, size_t workspace_bytes,
                    gnrf_stream_t stream);
/* include_vd=True (GazeNeRFNet(include_vd=True), models/gaze_nerf.py:29-30,70-80,140-141): RGB_layer_1's input becomes
 * [hidden 384 | view-direction encoding 27 | appearance 127], the encoding being the reference's Embedder (4 frequencies + input) of the
 * normalised ray direction.  The 27 columns are constant per RAY, so they fold into a per-ray bias of the fused kernel's last stage:
 *   gnrf_mlp_tc_pack_vd(params, n_vd = 27, ...)  packs a 192 x 538 RGB_layer_1 (n_vd = 0: same as gnrf_mlp_tc_pack);
 *   gnrf_mlp_tc_vd_bias(packed, ray_dl [B,N_r,4]) -> vd_bias [B,N_r,192];
 *   gnrf_mlp_tc_fwd_vd = gnrf_mlp_tc_fwd with vd_bias[i] per branch.  Inference only (the differentiable path has no view-direction
 *   operand; the drop-in module raises for grad with include_vd=True). */
int gnrf_mlp_tc_pack_vd(const float* const* params, int n_vd, void* packed, gnrf_stream_t stream);
int gnrf_mlp_tc_vd_bias(const void* packed, const float* ray_dl, int B, int N_r, float* vd_bias, gnrf_stream_t stream);
int gnrf_mlp_tc_fwd_vd(int n_branch, const void* const* packed, const float* const* bias, const float* const* vd_bias,
                       const float* ray_dl, const float* tvecs, const float* z_edges, int B, int N_r, int N_s, float* const* feat_ray,
                       float* const* bg_alpha, float* const* weights, void* workspace, size_t workspace_bytes, gnrf_stream_t stream);
/* Developer variant of gnrf_mlp_tc_fwd with explicit instrumentation arguments (no environment variables, no hidden state):
 * dbg_dump (nullable) [10][128][384] fp32 dump of tile 0's per-layer activations; timeline (nullable) [4][10][16] int64 clock64
 * stamps of CTA 0 (tests/tc_timeline.py); cluster_size 1 or 2 (CTAs sharing each weight stage by TMA multicast; 2 = default). */
int gnrf_mlp_tc_fwd_debug(int n_branch, const void* const* packed, const float* const* bias, const float* ray_dl,
                          const float* tvecs, const float* z_edges, int B, int N_r, int N_s, float* const* feat_ray,
                          float* const* bg_alpha, float* const* weights, void* workspace, size_t workspace_bytes, float* dbg_dump,
                          long long* timeline, int cluster_size, gnrf_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * Feature-map compose: background blend, gaze rotation of channel triplets, max-merge.
 * Replaces models/gaze_nerf.py:175-203 and rotate()/rotation_matrix_2d(), utils/model_utils.py:11-46.
 *   feat_* [B,C,P], a_* [B,P], bg [C,P], gaze [B,2] (rad)  ->  out [3,B,C,P] = (merge_face, eyes_planes, merge). */
int gnrf_compose_fwd(const float* feat_face, const float* a_face, const float* feat_eyes, const float* a_eyes,
                     const float* bg, const float* gaze, int B, int C, int P, float* out, gnrf_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * 2-D neural renderer (pixel-shuffle upsampler).  Replaces NeuralRenderer.forward (models/neural_renderer.py:
 * 98-113), PixelShuffleUpsample.forward and Blur.forward (models/pixel_shuffle_upsample.py:7-42).
 * params: HOST array of device pointers in state_dict order without bg_featmap and without the Blur buffers:
 *   for i < n_blocks: feat_upsample_list.i.layer_1 (w,b), layer_2 (w,b);
 *   then feat_2_rgb_list.0..n_blocks (w,b); then feat_layers.0..n_blocks-1 (w,b)     [n = 8*n_blocks + 2]
 *   featmap [N,C,S,S] -> img [N,3,S*2^n_blocks,S*2^n_blocks] in (0,1). */
size_t gnrf_nr_workspace_bytes(int N, int C, int S, int n_blocks, int min_feat);
int gnrf_neural_render_fwd(const float* const* params, const float* featmap, int N, int C, int S, int n_blocks,
                           int min_feat, float* img, void* workspace, size_t workspace_bytes, gnrf_stream_t stream);

/* Same computation on tcgen05 tensor cores (bf16x3 split precision, fp32 accumulate in TMEM).  gnrf_nr_tc_pack re-arranges the
 * weight matrices (hi/lo split, UMMA core-matrix K-slices) into `packed` (gnrf_nr_tc_packed_bytes, 128-byte aligned); redo after
 * every weight update.  params / workspace as for gnrf_neural_render_fwd.
 * gnrf_neural_render_tc_fwd runs the FUSED level kernels (csrc/nr_fused.cuh: per level, x = LeakyReLU(Blur(.)) formed on load -> W1 GEMM
 * (+ to-RGB rows) -> t1 in UMMA operand form; W2 GEMM -> drain -> W3 GEMM with the 4ci-wide pixel-shuffled tensor never leaving the
 * SM) whenever every level has a multiple of 128 pixels per image and fits the kernels' smem / TMEM budgets (true for the reference
 * configuration 258 ch, 64x64 -> 512x512), else the layer-wise path; gnrf_neural_render_tc_layerwise_fwd always runs the layer-wise
 * path (one conv_tc_kernel launch per 1x1 convolution, csrc/conv_tc.cu), kept as an on-device cross-check. */
size_t gnrf_nr_tc_packed_bytes(int C, int n_blocks, int min_feat);
int gnrf_nr_tc_pack(const float* const* params, int C, int n_blocks, int min_feat, void* packed, gnrf_stream_t stream);
int gnrf_neural_render_tc_fwd(const float* const* params, const void* packed, const float* featmap, int N, int C, int S,
                              int n_blocks, int min_feat, float* img, void* workspace, size_t workspace_bytes,
                              gnrf_stream_t stream);
int gnrf_neural_render_tc_layerwise_fwd(const float* const* params, const void* packed, const float* featmap, int N, int C, int S,
                                        int n_blocks, int min_feat, float* img, void* workspace, size_t workspace_bytes,
                                        gnrf_stream_t stream);

/* Multi-GPU batch sharding (one process per GPU): gnrf_neural_render_tc_fwd with the all-gather of the rendered images FUSED into
 * the kernel that produces them.  The first 3 * b_local images of the batch (merge_img_face | merge_img_eyes | merge_img of this
 * rank's faces) are also stored into every rank's gathered buffer [3][gb][3][P*P] -- through the NVSwitch multicast address
 * (multimem.st, NVLS) when mc_ptr != NULL, else with one store per peer over NVLink.  peer_ptrs: HOST array of `world` device
 * pointers (each rank's symmetric buffer mapped into this process).  Cross-rank ordering (buffer reuse, consumption) is the caller's:
 * gazenerf_b200/dist.py double-buffers and runs one device barrier per step.  Replaces the all-gather a multi-GPU caller of
 * GazeNeRFNet.forward would issue (the reference is single-GPU, trainer/base.py:31-35). */
int gnrf_neural_render_tc_fwd_gather(const float* const* params, const void* packed, const float* featmap, int N, int C, int S,
                                     int n_blocks, int min_feat, float* img, void* workspace, size_t workspace_bytes,
                                     float* const* peer_ptrs, float* mc_ptr, int world, int rank, int b_local, int gb,
                                     gnrf_stream_t stream);

/* =========================================================================================================
 * Training path (forward that keeps activations + backward).  The reference gets its gradients from torch
 * autograd over the same graph (train.py -> trainer/gazenerf_trainer.py:479-528, loss.backward()); these entry
 * points are the stages gazenerf_b200/train.py chains inside ONE torch.autograd.Function.  Everything is
 * channel-major: [image][channel][pixel or sample point], points contiguous; "img stride" arguments are the
 * element distance between consecutive images (0 = dense).
 * ========================================================================================================= */
enum { GNRF_ACT_NONE = 0, GNRF_ACT_RELU = 1, GNRF_ACT_LRELU02 = 2 };

/* Generic 1x1 convolution / per-point Linear on tcgen05 (bf16x3 split, fp32 accumulate):
 *   out[img][n][p] = act( sum_k W[n][k] X[img][k][p] + bias[n] + bias_img[img][n] ) * (mask > 0 ? 1 : mask_slope) + add
 * Replaces nn.Conv2d(k=1) forward (models/mlp_nerf.py:29-93, models/pixel_shuffle_upsample.py:26-31) and, with the
 * transposed weight packed, its input gradient.  mask / add / bias_img may be NULL; mask applies to rows < mask_rows and
 * add to rows < add_rows (0 = all rows).  gnrf_conv_tc_pack: W dense [N][K], or with transposed != 0 a dense [K][N] matrix
 * whose transpose is used; bias may be NULL.  packed: gnrf_conv_tc_packed_bytes(N, K), 128-byte aligned. */
size_t gnrf_conv_tc_packed_bytes(int N, int K);
int gnrf_conv_tc_pack(const float* W, const float* bias, int N, int K, int transposed, void* packed, gnrf_stream_t stream);
int gnrf_conv_tc(const void* packed, int N, int K, const float* X, long long x_img_stride, const float* bias_img, float* out,
                 long long out_img_stride, int act, const float* mask, long long mask_img_stride, int mask_rows,
                 float mask_slope, const float* add, long long add_img_stride, int add_rows, int n_img, int HW,
                 gnrf_stream_t stream);

/* Weight gradient of the same layer on tcgen05:  dW[n][k] (=, += if accumulate) sum_img sum_p dY[img][n][p] X[img][k][p];
 * db (nullable): db_sum == 0 -> [n_img][N] per-image sum_p dY (the per-face folded-code bias), else [N] summed over images.
 * Deterministic split-K (partials in `workspace`, gnrf_wgrad_tc_workspace_bytes).  HW and strides multiples of 4. */
size_t gnrf_wgrad_tc_workspace_bytes(int N, int K, int n_img, int HW);
int gnrf_wgrad_tc(const float* dY, long long dy_img_stride, const float* X, long long x_img_stride, int N, int K, int n_img,
                  int HW, float* dW, float* db, int db_sum, int accumulate, void* workspace, size_t workspace_bytes,
                  gnrf_stream_t stream);

/* Positional encoding of the sample points, channel-major (Embedder.forward, utils/model_utils.py:272-280, on
 * pts = o + d*l*z, :315):  pe [B][63][N_r*N_s].  gnrf_pe_bwd: gradients arriving at the encoding from layer 0 (g_pe_a) and
 * from the skip layer (g_pe_b, nullable) -> per-ray  g_m [B][N_r][3] (d*l), g_o [B][N_r][3], g_z [B][N_r][N_s+1]  (all +=). */
int gnrf_pe_fwd(const float* ray_dl, const float* tvecs, const float* z_edges, int B, int N_r, int N_s, float* pe,
                long long pe_img_stride, gnrf_stream_t stream);
int gnrf_pe_bwd(const float* g_pe_a, long long ga_stride, const float* g_pe_b, long long gb_stride, const float* pe,
                long long pe_stride, const float* ray_dl, const float* z_edges, int B, int N_r, int N_s, float* g_m, float* g_o,
                float* g_z, gnrf_stream_t stream);

/* Alpha compositing on channel-major activations (CalcRayColor.forward, utils/model_utils.py:493-534; sigma = ReLU(raw)):
 *   h [B][C][N_r*N_s], sigma_raw [B][N_r*N_s]  ->  Hc [B][C+1][N_r] (row C = sum_k w_k), bg_alpha [B][N_r], weights [B][N_r][N_s].
 * Backward: g_Hc [B][C+1][N_r], g_bg_alpha (nullable) -> g_h [B][C][P] (already multiplied by h > 0: h is a post-ReLU
 * activation), g_sigma [B][P] (w.r.t. the raw density), g_z [B][N_r][N_s+1] (+=), g_l [B][N_r] (+=). */
int gnrf_composite_cm_fwd(const float* h, long long h_stride, const float* sigma_raw, long long s_stride, const float* z_edges,
                          const float* ray_dl, int B, int N_r, int N_s, int C, float* Hc, float* bg_alpha, float* weights,
                          gnrf_stream_t stream);
int gnrf_composite_cm_bwd(const float* g_Hc, const float* g_bg_alpha, const float* h, long long h_stride, const float* sigma_raw,
                          long long s_stride, const float* weights, const float* z_edges, const float* ray_dl, int B, int N_r,
                          int N_s, int C, float* g_h, long long gh_stride, float* g_sigma, long long gs_stride, float* g_z,
                          float* g_l, gnrf_stream_t stream);

/* ---- The same layers on PRE-SPLIT bf16 planes (csrc/lin_hl.cu; training path of models/mlp_nerf.py:95-119 without a conversion
 * pass).  A plane tensor holds x as hi = bf16(x) [and lo = bf16(x - hi), `plane_stride` elements after hi]: [planes][img][rows][HW],
 * points contiguous; planes = 2 is the bf16x3 scheme of the kernels above, planes = 1 single-pass bf16 (half the bytes; gradient
 * tolerance stated in tests/test_train_grad.py).  All strides in ELEMENTS (bf16 or float); HW a multiple of 256, K >= 32.
 *   gnrf_lin_hl:  y[img][n][p] = act( sum_k W[n][k] x[img][k][p] + bias[n] + bias_img[img][n] ) * mask_bit[img][n][p] (n < mask_rows)
 *     act (GNRF_ACT_NONE / GNRF_ACT_RELU) applies to rows n < act_rows (0 = all rows; a multiple of 32 otherwise);
 *     rows n < hl_rows (a multiple of 32, or >= N) are written as planes to `out`, rows n >= hl_rows as fp32 to
 *     out_f32[img][n - hl_rows][p] (for the non-GEMM consumers: composite, positional-encoding backward).
 *     mask_out (nullable): uint32 [img][rows][HW/32], bit (p % 32) of word p/32 = (y > 0) -- the ReLU sign bits the input gradient of
 *     the NEXT layer needs; mask_bits (nullable): such a tensor, applied to this call's output (input-gradient use).  Mask strides
 *     in 32-bit words.
 *   gnrf_wgrad_hl: dW[n][k] = sum_img sum_p dY[img][n][p] X[img][k][p]; db as in gnrf_wgrad_tc.  N >= 128, HW a multiple of 32.
 *   gnrf_pe_fwd_hl: gnrf_pe_fwd that also writes the encoding as planes; gnrf_composite_cm_bwd_hl: g_h / g_sigma as planes. */
size_t gnrf_lin_hl_packed_bytes(int N, int K, int planes);
int gnrf_lin_hl_pack(const float* W, const float* bias, int N, int K, int transposed, int planes, void* packed, gnrf_stream_t stream);
int gnrf_lin_hl(const void* packed, int N, int K, int planes, const void* X, long long x_img_stride, long long x_plane_stride,
                const float* bias_img, int act, int act_rows, void* out, long long out_img_stride, long long out_plane_stride, int hl_rows,
                float* out_f32, long long f32_img_stride, const void* mask_bits, long long mask_img_stride, int mask_rows,
                void* mask_out, long long mask_out_img_stride, int mask_out_rows, int n_img, int HW, gnrf_stream_t stream);
size_t gnrf_wgrad_hl_workspace_bytes(int N, int K, int n_img, int HW);
int gnrf_wgrad_hl(const void* dY, long long dy_img_stride, long long dy_plane_stride, const void* X, long long x_img_stride,
                  long long x_plane_stride, int planes, int N, int K, int n_img, int HW, float* dW, float* db, int db_sum,
                  void* workspace, size_t workspace_bytes, gnrf_stream_t stream);
int gnrf_pe_fwd_hl(const float* ray_dl, const float* tvecs, const float* z_edges, int B, int N_r, int N_s, float* pe,
                   long long pe_img_stride, void* hl, long long hl_img_stride, long long hl_plane_stride, int planes,
                   gnrf_stream_t stream);
int gnrf_composite_cm_bwd_hl(const float* g_Hc, const float* g_bg_alpha, const float* h, long long h_stride, const float* sigma_raw,
                             long long s_stride, const float* weights, const float* z_edges, const float* ray_dl, int B, int N_r,
                             int N_s, int C, void* g_h, long long gh_stride, long long gh_plane_stride, void* g_sigma,
                             long long gs_stride, long long gs_plane_stride, int planes, float* g_z, float* g_l,
                             gnrf_stream_t stream);

/* Ray geometry backward (GenSamplePoints, utils/model_utils.py:364-372): per-ray gradients -> contrib [B][N_r][12] =
 * (dL/dR 3x3 row-major, dL/dT 3); the caller sums over rays. */
int gnrf_geom_bwd(const float* xy, const float* rmats, const float* inv_inmats, const float* g_m, const float* g_o,
                  const float* g_l, const float* g_z, int B, int N_r, int N_s, float* contrib, gnrf_stream_t stream);

/* Backward of gnrf_compose_fwd.  g_out [3][B][C][P] -> g_feat_* [B][C][P], g_bg [C][P], and per channel-group partial sums the caller
 * adds up in order: g_a_* [gnrf_compose_bwd_groups(C)][B][P], g_gaze_part [B][gnrf_compose_bwd_blocks(P, C)][2]. */
int gnrf_compose_bwd_groups(int C);
int gnrf_compose_bwd_blocks(int P, int C);
int gnrf_compose_bwd(const float* g_out, const float* feat_face, const float* a_face, const float* feat_eyes, const float* a_eyes,
                     const float* bg, const float* gaze, int B, int C, int P, float* g_feat_face, float* g_a_face,
                     float* g_feat_eyes, float* g_a_eyes, float* g_bg, float* g_gaze_part, gnrf_stream_t stream);

/* Neural renderer, training: forward that keeps per-level activations in `saved` (gnrf_nr_train_saved_bytes), and its
 * backward -> g_featmap [N][C][S][S] and g_params (HOST array of device pointers, same order / shapes as params). */
size_t gnrf_nr_train_saved_bytes(int N, int C, int S, int n_blocks, int min_feat);
int gnrf_nr_train_fwd(const float* const* params, const void* packed, const float* featmap, int N, int C, int S, int n_blocks,
                      int min_feat, float* img, void* saved, size_t saved_bytes, gnrf_stream_t stream);
size_t gnrf_nr_train_bwd_workspace_bytes(int N, int C, int S, int n_blocks, int min_feat);
int gnrf_nr_train_bwd(const float* const* params, const float* featmap, const void* saved, const float* img, const float* g_img,
                      int N, int C, int S, int n_blocks, int min_feat, float* g_featmap, float* const* g_params, void* workspace,
                      size_t workspace_bytes, gnrf_stream_t stream);

/* Fused data terms of GazeNeRFLoss (losses/gazenerf_loss.py:294-352; mask algebra of calc_total_loss :420-424).
 * images [B,3,HW] (bg_img [1,3,HW]), masks [B,1,HW] float.  terms[5] = (head, eyes, face, nonhead, bg), each the mean the reference
 * takes over its boolean gather (an empty mask yields NaN, as there); sums[9] keeps the numerators / pixel counts for the backward.
 * workspace: gnrf_data_loss_workspace_floats() floats.  Backward: g_terms[5] -> gradients of the four images. */
size_t gnrf_data_loss_workspace_floats(void);
int gnrf_data_loss_fwd(const float* img_face, const float* img_eyes, const float* img, const float* bg_img, const float* gt,
                       const float* face_mask, const float* full_eye, const float* left_eye, const float* right_eye, int B, int HW,
                       int use_l1, float bg_value, float* terms, float* sums, float* workspace, gnrf_stream_t stream);
int gnrf_data_loss_bwd(const float* img_face, const float* img_eyes, const float* img, const float* bg_img, const float* gt,
                       const float* face_mask, const float* full_eye, const float* left_eye, const float* right_eye, int B, int HW,
                       int use_l1, float bg_value, const float* sums, const float* g_terms, float* g_img_face, float* g_img_eyes,
                       float* g_img, float* g_bg_img, gnrf_stream_t stream);

/* =========================================================================================================
 * Dataset sample -> device tensors (the step before the path).  Replaces the per-item transforms of
 * GazeDataset.__getitem__ (datasets/eth_xgaze.py:308-360) and GazeNeRFTrainer.prepare_data (trainer/gazenerf_trainer.py:250-337)
 * on the raw HDF5 records (schema: dataset_pre_processing.py:260-380).  All pointers are DEVICE pointers to the raw u8 / f64 arrays
 * (the caller uploads the records as stored: 0.98 MB per sample instead of 6 MB of fp32 tensors).
 *   face_patch_bgr u8 [B,H,W,3] -> img f32 [B,3,H,W] = RGB / 255;   head_mask u8 [B,H,W] -> cv2.erode(3x3 ones, erode_iterations)
 *   -> f32 [B,1,H,W] (mask VALUES are kept, e.g. 0 / 255);   left / right eye masks -> f32 [B,1,H,W].   W % 4 == 0.
 * ========================================================================================================= */
int gnrf_sample_images_to_device(const uint8_t* face_patch_bgr, const uint8_t* head_mask, const uint8_t* left_eye_mask,
                                 const uint8_t* right_eye_mask, int B, int H, int W, int erode_iterations, float* img,
                                 float* head_mask_f, float* left_eye_mask_f, float* right_eye_mask_f, gnrf_stream_t stream);
/* code_row0 f64 [306] (row 0 of the subject's latent_codes), code_rows f64 [B,306] (the samples' own rows; only [279:] = illumination
 * is taken from them, eth_xgaze.py:346-347), pitchyaw f64 [B,2], c2w_Rmat f64 [B,3,3], c2w_Tvec f64 [B,3], inmat f64 [B,3,3]
 *   -> iden f32 [B,100], expr [B,79], text [B,100], illu [B,27], gaze [B,2], rmats [B,3,3], tvecs [B,3,1],
 *      inv_inmats [B,3,3] = inverse of inmat with rows 0,1 scaled by featmap_size / img_size (f64 arithmetic, then cast). */
int gnrf_sample_meta_to_device(const double* code_row0, const double* code_rows, const double* pitchyaw, const double* c2w_Rmat,
                               const double* c2w_Tvec, const double* inmat, int B, int featmap_size, int img_size, float* iden,
                               float* expr, float* text, float* illu, float* gaze, float* rmats, float* tvecs, float* inv_inmats,
                               gnrf_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* GNRF_H_ */
