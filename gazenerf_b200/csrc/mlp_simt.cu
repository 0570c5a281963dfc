// Exact-fp32 CUDA-core radiance MLP (one fused kernel per branch) + alpha compositing.
//
// This is the literal path: it evaluates the reference graph as written (244-wide input with the per-face codes
// broadcast to every point, skip concat, 511-wide RGB_layer_1, 258-wide per-point output), with no algebraic
// folds and fp32 FMA arithmetic only.  It serves (a) sample counts the tensor-core kernel is not specialised
// for and (b) as the on-device cross-check of the folded bf16x3 tcgen05 kernel at full size.
//
// Reference: utils/model_utils.py:272-280 (Embedder), models/gaze_nerf.py:248-262,136-143 (code concat),
//            models/mlp_nerf.py:95-119 (MLPforNeRF.forward), utils/model_utils.py:493-534 (CalcRayColor).
#include "common.cuh"

namespace gnrf {

constexpr int kTP = 64;        // points per CTA tile
constexpr int kThreads = 256;  // 4 point-groups (16 pts) x 64 column lanes (6 columns each, stride 64)
constexpr int kVpDim = GNRF_PE_DIMS + GNRF_SHAPE_EXT_DIMS;  // 244
constexpr int kHidMax = 384;
constexpr int kKC = 8;         // K chunk staged through smem
constexpr int kWLd = 388;      // smem leading dim of the staged weight chunk (== 4 mod 32: conflict-free fill)
constexpr int kColsPerThread = 6;
constexpr int kPtsPerThread = 16;

struct MlpParams {
  const float* w[12];
  const float* b[12];
};

struct SimtSmem {
  float vp[kTP][kVpDim];        // [PE 63 | shape_ext 181]
  float hid[kTP][kHidMax];      // current hidden activations (updated in place through registers)
  float wch[2][kKC][kWLd];      // double-buffered weight chunk, transposed: [k][n]
  float appea[GNRF_APPEA_DIMS];
  float cvec[kHidMax];          // per-face constant part of RGB_layer_1 (appearance columns)
  float sigma[kTP];
};

// acc[pt][j] += sum_k src[p][k] * W[n][wcol0 + k],  n = lane_n + 64 j.
__device__ __forceinline__ void accumulate(float (&acc)[kPtsPerThread][kColsPerThread], SimtSmem& sm, const float* src,
                                           int src_ld, int K, const float* __restrict__ W, int ldw, int wcol0, int N) {
  const int tid = threadIdx.x;
  const int lane_n = tid & 63;
  const int pg = tid >> 6;
  const int n_chunks = (K + kKC - 1) / kKC;
  constexpr int kLoadsPerThread = (kHidMax * kKC) / kThreads;  // 12
  float stage[kLoadsPerThread];

  auto fetch = [&](int chunk) {
    int k0 = chunk * kKC;
#pragma unroll
    for (int i = 0; i < kLoadsPerThread; ++i) {
      int idx = tid + i * kThreads;
      int n = idx >> 3, kk = idx & 7;
      float v = 0.0f;
      if (n < N && k0 + kk < K) v = __ldg(W + (size_t)n * ldw + wcol0 + k0 + kk);
      stage[i] = v;
    }
  };
  auto commit = [&](int buf) {
#pragma unroll
    for (int i = 0; i < kLoadsPerThread; ++i) {
      int idx = tid + i * kThreads;
      sm.wch[buf][idx & 7][idx >> 3] = stage[i];
    }
  };

  fetch(0);
  commit(0);
  __syncthreads();
  for (int c = 0; c < n_chunks; ++c) {
    const int buf = c & 1;
    if (c + 1 < n_chunks) fetch(c + 1);
    const int k0 = c * kKC;
    const int klim = min(kKC, K - k0);
    if (klim == kKC) {
#pragma unroll
      for (int kk = 0; kk < kKC; ++kk) {
        float wv[kColsPerThread];
#pragma unroll
        for (int j = 0; j < kColsPerThread; ++j) wv[j] = sm.wch[buf][kk][lane_n + 64 * j];
#pragma unroll
        for (int p = 0; p < kPtsPerThread; ++p) {
          float a = src[(pg * kPtsPerThread + p) * src_ld + k0 + kk];
#pragma unroll
          for (int j = 0; j < kColsPerThread; ++j) acc[p][j] = fmaf(a, wv[j], acc[p][j]);
        }
      }
    } else {
      for (int kk = 0; kk < klim; ++kk) {
        float wv[kColsPerThread];
#pragma unroll
        for (int j = 0; j < kColsPerThread; ++j) wv[j] = sm.wch[buf][kk][lane_n + 64 * j];
#pragma unroll
        for (int p = 0; p < kPtsPerThread; ++p) {
          float a = src[(pg * kPtsPerThread + p) * src_ld + k0 + kk];
#pragma unroll
          for (int j = 0; j < kColsPerThread; ++j) acc[p][j] = fmaf(a, wv[j], acc[p][j]);
        }
      }
    }
    if (c + 1 < n_chunks) commit(buf ^ 1);
    __syncthreads();
  }
}

__device__ __forceinline__ void zero_acc(float (&acc)[kPtsPerThread][kColsPerThread]) {
#pragma unroll
  for (int p = 0; p < kPtsPerThread; ++p)
#pragma unroll
    for (int j = 0; j < kColsPerThread; ++j) acc[p][j] = 0.0f;
}

// hid[p][n] = act(acc + bias[n] (+ extra[n]))
__device__ __forceinline__ void store_hidden(float (&acc)[kPtsPerThread][kColsPerThread], SimtSmem& sm,
                                             const float* __restrict__ bias, const float* extra, int N, bool relu) {
  const int lane_n = threadIdx.x & 63;
  const int pg = threadIdx.x >> 6;
#pragma unroll
  for (int j = 0; j < kColsPerThread; ++j) {
    int n = lane_n + 64 * j;
    if (n < N) {
      float bv = __ldg(bias + n);
      if (extra != nullptr) bv += extra[n];
#pragma unroll
      for (int p = 0; p < kPtsPerThread; ++p) {
        float v = acc[p][j] + bv;
        if (relu) v = fmaxf(v, 0.0f);
        sm.hid[pg * kPtsPerThread + p][n] = v;
      }
    }
  }
}

__global__ void __launch_bounds__(kThreads, 1)
mlp_simt_kernel(MlpParams prm, const float4* __restrict__ ray_dl, const float* __restrict__ tvecs,
                const float* __restrict__ z_edges, const float* __restrict__ shape_ext, const float* __restrict__ appea,
                int N_r, int N_s, int hidden, int n_feat, int tiles_per_face, float* __restrict__ feat_pts,
                float* __restrict__ sigma_pts) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SimtSmem& sm = *reinterpret_cast<SimtSmem*>(smem_raw);
  const int tid = threadIdx.x;
  const int b = blockIdx.x / tiles_per_face;
  const int tile = blockIdx.x - b * tiles_per_face;
  const int pts_per_face = N_r * N_s;
  const int p0 = tile * kTP;

  // ---- prologue: sample positions -> positional encoding; broadcast codes --------------------------------
  if (tid < kTP * 3) {
    int p = tid / 3, c = tid - p * 3;
    int gp = min(p0 + p, pts_per_face - 1);
    int ray = gp / N_s, s = gp - ray * N_s;
    float4 dl = ray_dl[b * N_r + ray];
    float dc = (c == 0) ? dl.x : ((c == 1) ? dl.y : dl.z);
    float z = z_edges[((size_t)b * N_r + ray) * (N_s + 1) + s];
    // pts = o + ((d * l) * z)   (utils/model_utils.py:315)
    float x = __fadd_rn(tvecs[b * 3 + c], __fmul_rn(__fmul_rn(dc, dl.w), z));
    sm.vp[p][c] = x;
    float f = 1.0f;
#pragma unroll
    for (int q = 0; q < 10; ++q) {
      float sv, cv;
      sincosf(__fmul_rn(x, f), &sv, &cv);
      sm.vp[p][3 + 6 * q + c] = sv;
      sm.vp[p][6 + 6 * q + c] = cv;
      f *= 2.0f;
    }
  }
  for (int i = tid; i < kTP * GNRF_SHAPE_EXT_DIMS; i += kThreads) {
    int p = i / GNRF_SHAPE_EXT_DIMS, c = i - p * GNRF_SHAPE_EXT_DIMS;
    sm.vp[p][GNRF_PE_DIMS + c] = shape_ext[b * GNRF_SHAPE_EXT_DIMS + c];
  }
  if (tid < GNRF_APPEA_DIMS) sm.appea[tid] = appea[b * GNRF_APPEA_DIMS + tid];
  __syncthreads();
  // per-face constant of RGB_layer_1: cvec[n] = sum_k appea[k] * W_rgb1[n][hidden + k]
  const int h2 = hidden / 2;
  for (int n = tid; n < h2; n += kThreads) {
    const float* wr = prm.w[10] + (size_t)n * (hidden + GNRF_APPEA_DIMS) + hidden;
    float s = 0.0f;
    for (int k = 0; k < GNRF_APPEA_DIMS; ++k) s = fmaf(sm.appea[k], __ldg(wr + k), s);
    sm.cvec[n] = s;
  }
  __syncthreads();

  float acc[kPtsPerThread][kColsPerThread];

  // ---- FeaExt_module_0..7 with skip concat after layer 4 (models/mlp_nerf.py:101-107) -------------------
  zero_acc(acc);
  accumulate(acc, sm, &sm.vp[0][0], kVpDim, kVpDim, prm.w[0], kVpDim, 0, hidden);
  store_hidden(acc, sm, prm.b[0], nullptr, hidden, true);
  __syncthreads();
  for (int layer = 1; layer < 8; ++layer) {
    zero_acc(acc);
    if (layer == 5) {
      const int ldw = kVpDim + hidden;
      accumulate(acc, sm, &sm.vp[0][0], kVpDim, kVpDim, prm.w[5], ldw, 0, hidden);
      accumulate(acc, sm, &sm.hid[0][0], kHidMax, hidden, prm.w[5], ldw, kVpDim, hidden);
    } else {
      accumulate(acc, sm, &sm.hid[0][0], kHidMax, hidden, prm.w[layer], hidden, 0, hidden);
    }
    // accumulate() ends with __syncthreads(): every read of hid is complete before it is overwritten
    store_hidden(acc, sm, prm.b[layer], nullptr, hidden, true);
    __syncthreads();
  }

  // ---- density head: sigma = ReLU(w_d . x + b_d) (models/mlp_nerf.py:109,115) ----------------------------
  {
    int p = tid >> 2, part = tid & 3;
    float s = 0.0f;
    for (int k = part; k < hidden; k += 4) s = fmaf(sm.hid[p][k], __ldg(prm.w[8] + k), s);
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    if (part == 0) sm.sigma[p] = fmaxf(s + __ldg(prm.b[8]), 0.0f);
  }
  __syncthreads();

  // ---- RGB_layer_0 (no activation), RGB_layer_1 (+appearance, ReLU), RGB_layer_2 (:110-113) ---------------
  zero_acc(acc);
  accumulate(acc, sm, &sm.hid[0][0], kHidMax, hidden, prm.w[9], hidden, 0, hidden);
  store_hidden(acc, sm, prm.b[9], nullptr, hidden, false);
  __syncthreads();
  zero_acc(acc);
  accumulate(acc, sm, &sm.hid[0][0], kHidMax, hidden, prm.w[10], hidden + GNRF_APPEA_DIMS, 0, h2);
  store_hidden(acc, sm, prm.b[10], sm.cvec, h2, true);
  __syncthreads();
  zero_acc(acc);
  accumulate(acc, sm, &sm.hid[0][0], kHidMax, h2, prm.w[11], h2, 0, n_feat);

  // ---- write per-point outputs ------------------------------------------------------------------------
  {
    const int lane_n = tid & 63, pg = tid >> 6;
#pragma unroll
    for (int j = 0; j < kColsPerThread; ++j) {
      int n = lane_n + 64 * j;
      if (n < n_feat) {
        float bv = __ldg(prm.b[11] + n);
#pragma unroll
        for (int p = 0; p < kPtsPerThread; ++p) {
          int gp = p0 + pg * kPtsPerThread + p;
          if (gp < pts_per_face) feat_pts[((size_t)b * pts_per_face + gp) * n_feat + n] = acc[p][j] + bv;
        }
      }
    }
    if (tid < kTP && p0 + tid < pts_per_face) sigma_pts[(size_t)b * pts_per_face + p0 + tid] = sm.sigma[tid];
  }
}

// One CTA per ray: weights by the sequential cumprod recurrence, then the weighted channel sums.
__global__ void __launch_bounds__(256)
composite_kernel(const float* __restrict__ feat_pts, const float* __restrict__ sigma_pts, const float* __restrict__ z_edges,
                 const float4* __restrict__ ray_dl, int N_r, int N_s, int n_feat, float* __restrict__ feat_ray,
                 float* __restrict__ bg_alpha, float* __restrict__ depth, float* __restrict__ weights) {
  extern __shared__ float s_w[];  // [N_s] alpha then weights
  const int ray = blockIdx.x;     // flattened (b, r)
  const int b = ray / N_r, r = ray - b * N_r;
  const int tid = threadIdx.x;
  const float* ze = z_edges + (size_t)ray * (N_s + 1);
  const float l = ray_dl[ray].w;
  for (int k = tid; k < N_s; k += blockDim.x) {
    float delta = __fmul_rn(__fsub_rn(ze[k + 1], ze[k]), l);
    float sg = sigma_pts[(size_t)ray * N_s + k];
    s_w[k] = __fsub_rn(1.0f, expf(-__fmul_rn(sg, delta)));  // alpha (utils/model_utils.py:500)
  }
  __syncthreads();
  if (tid == 0) {
    float T = 1.0f, acc_w = 0.0f, acc_d = 0.0f;
    for (int k = 0; k < N_s; ++k) {
      float a = s_w[k];
      float wk = __fmul_rn(a, T);  // w_k = alpha_k * T_k (:512)
      T = __fmul_rn(T, __fadd_rn(__fsub_rn(1.0f, a), 1e-10f));  // T_{k+1} = T_k (1 - alpha_k + 1e-10) (:508-510)
      s_w[k] = wk;
      acc_w += wk;
      acc_d = fmaf(wk, ze[k], acc_d);
    }
    bg_alpha[ray] = 1.0f - acc_w;
    if (depth != nullptr) depth[ray] = acc_d;
  }
  __syncthreads();
  if (weights != nullptr)
    for (int k = tid; k < N_s; k += blockDim.x) weights[(size_t)ray * N_s + k] = s_w[k];
  const float* f = feat_pts + (size_t)ray * N_s * n_feat;
  for (int c = tid; c < n_feat; c += blockDim.x) {
    float acc = 0.0f;
    for (int k = 0; k < N_s; ++k) acc = fmaf(s_w[k], f[(size_t)k * n_feat + c], acc);
    feat_ray[((size_t)b * n_feat + c) * N_r + r] = acc;
  }
}

}  // namespace gnrf

using namespace gnrf;

extern "C" int gnrf_mlp_simt_fwd(const float* const* params, const float* ray_dl, const float* tvecs, const float* z_edges,
                                 const float* shape_ext, const float* appea, int B, int N_r, int N_s, int hidden, int n_feat,
                                 float* feat_pts, float* sigma_pts, gnrf_stream_t stream) {
  GNRF_CHECK_ARG(params && ray_dl && tvecs && z_edges && shape_ext && appea && feat_pts && sigma_pts);
  GNRF_CHECK_ARG(B > 0 && N_r > 0 && N_s > 0);
  if (hidden > kHidMax || hidden % 2 != 0 || n_feat > kHidMax || n_feat < 1)
    return fail(GNRF_ERR_UNSUPPORTED, "gnrf_mlp_simt_fwd: hidden=%d n_feat=%d (need even hidden <= %d, n_feat <= %d)", hidden,
                n_feat, kHidMax, kHidMax);
  MlpParams prm;
  for (int i = 0; i < 12; ++i) {
    prm.w[i] = params[2 * i];
    prm.b[i] = params[2 * i + 1];
    GNRF_CHECK_ARG(prm.w[i] != nullptr && prm.b[i] != nullptr);
  }
  {
    int n_sm = 0;
    int rc = device_once(kOnceMlpSimt, &n_sm, []() -> int {
      GNRF_CUDA(cudaFuncSetAttribute(mlp_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SimtSmem)));
      return GNRF_OK;
    });
    if (rc != GNRF_OK) return rc;
  }
  int tiles_per_face = ceil_div(N_r * N_s, kTP);
  mlp_simt_kernel<<<B * tiles_per_face, kThreads, sizeof(SimtSmem), as_stream(stream)>>>(
      prm, reinterpret_cast<const float4*>(ray_dl), tvecs, z_edges, shape_ext, appea, N_r, N_s, hidden, n_feat, tiles_per_face,
      feat_pts, sigma_pts);
  GNRF_LAUNCH_CHECK();
  count_launches(1);
  return GNRF_OK;
}

extern "C" int gnrf_composite_fwd(const float* feat_pts, const float* sigma_pts, const float* z_edges, const float* ray_dl,
                                  int B, int N_r, int N_s, int n_feat, float* feat_ray, float* bg_alpha, float* depth,
                                  float* weights, gnrf_stream_t stream) {
  GNRF_CHECK_ARG(feat_pts && sigma_pts && z_edges && ray_dl && feat_ray && bg_alpha);
  GNRF_CHECK_ARG(B > 0 && N_r > 0 && N_s > 0 && n_feat > 0 && N_s <= 8192);
  composite_kernel<<<B * N_r, 256, N_s * sizeof(float), as_stream(stream)>>>(
      feat_pts, sigma_pts, z_edges, reinterpret_cast<const float4*>(ray_dl), N_r, N_s, n_feat, feat_ray, bg_alpha, depth, weights);
  GNRF_LAUNCH_CHECK();
  count_launches(1);
  return GNRF_OK;
}
