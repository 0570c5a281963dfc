// 2-D neural renderer: 1x1-conv GEMMs with fused epilogues, pixel-shuffle, 3x3 binomial blur, RGB skip pyramid.
// Reference: models/neural_renderer.py:98-113 (NeuralRenderer.forward),
//            models/pixel_shuffle_upsample.py:7-42 (Blur -> kornia.filters.filter2d, PixelShuffleUpsample).
#include "common.cuh"
#include "conv_tc.cuh"
#include "nr_plan.cuh"
#include "nr_fused.cuh"

namespace gnrf {

constexpr int BM = 64;    // output channels per CTA
constexpr int BN = 128;   // pixels per CTA
constexpr int BK = 8;
constexpr int kGemmThreads = 256;

enum Epilogue { EPI_LRELU = 0, EPI_PSU = 1 };

__device__ __forceinline__ float lrelu(float v) { return v >= 0.0f ? v : 0.2f * v; }

// out[n][co][p] = epi( sum_ci W[co][ci] * X[n][ci][p] + bias[co] ),  X: [N, K, HW] (NCHW, 1x1 conv == GEMM).
// EPI_PSU: v = lrelu(.) + res[n][co % Cres][p], written pixel-shuffled (factor 2):
//          out[n][co/4][2h + (co%4)/2][2w + (co%4)%2]   (pixel_shuffle_upsample.py:34-40)
template <int EPI>
__global__ void __launch_bounds__(kGemmThreads)
conv1x1_kernel(const float* __restrict__ W, const float* __restrict__ bias, const float* __restrict__ X, int M, int K, int HW,
               int Wd, const float* __restrict__ res, int Cres, float* __restrict__ out) {
  __shared__ __align__(16) float As[2][BK][BM];
  __shared__ __align__(16) float Bs[2][BK][BN];
  const int tid = threadIdx.x;
  const int n = blockIdx.z;
  const int co0 = blockIdx.y * BM;
  const int p0 = blockIdx.x * BN;
  const float* Xn = X + (size_t)n * K * HW;
  const int ty = tid >> 4, tx = tid & 15;  // thread tile: 4 co x (4 + 4) px

  // global->register staging
  float a_st[2];
  float4 b_st;
  auto fetch = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      int idx = tid + i * kGemmThreads;  // 0..511
      int m = idx >> 3, kk = idx & 7;
      a_st[i] = (co0 + m < M && k0 + kk < K) ? __ldg(W + (size_t)(co0 + m) * K + k0 + kk) : 0.0f;
    }
    int kk = tid >> 5, c4 = tid & 31;
    int px = p0 + c4 * 4;
    b_st = (k0 + kk < K && px < HW) ? __ldg(reinterpret_cast<const float4*>(Xn + (size_t)(k0 + kk) * HW + px))
                                    : make_float4(0.f, 0.f, 0.f, 0.f);
  };
  auto commit = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      int idx = tid + i * kGemmThreads;
      As[buf][idx & 7][idx >> 3] = a_st[i];
    }
    int kk = tid >> 5, c4 = tid & 31;
    *reinterpret_cast<float4*>(&Bs[buf][kk][c4 * 4]) = b_st;
  };

  float acc[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;

  const int n_k = (K + BK - 1) / BK;
  fetch(0);
  commit(0);
  __syncthreads();
  for (int kc = 0; kc < n_k; ++kc) {
    const int buf = kc & 1;
    if (kc + 1 < n_k) fetch((kc + 1) * BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float4 a = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
      float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
      float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][64 + tx * 4]);
      float av[4] = {a.x, a.y, a.z, a.w};
      float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (kc + 1 < n_k) commit(buf ^ 1);
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int co = co0 + ty * 4 + i;
    if (co >= M) continue;
    float bv = __ldg(bias + co);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int px = p0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (px >= HW) continue;
      float v = lrelu(acc[i][j] + bv);
      if (EPI == EPI_LRELU) {
        out[((size_t)n * M + co) * HW + px] = v;
      } else {
        v += __ldg(res + ((size_t)n * Cres + (co % Cres)) * HW + px);
        int h = px / Wd, w = px - h * Wd;
        int c = co >> 2, si = (co >> 1) & 1, sj = co & 1;
        out[(((size_t)n * (M >> 2) + c) * (2 * (HW / Wd)) + (2 * h + si)) * (2 * Wd) + 2 * w + sj] = v;
      }
    }
  }
}

__device__ __forceinline__ int reflect(int i, int n) { return i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i); }

// Depthwise [1,2,1]x[1,2,1]/16 with reflect border (pixel_shuffle_upsample.py:7-16). planes = N*C.
__global__ void blur3x3_kernel(const float* __restrict__ in, int H, int Wd, long long total, float* __restrict__ out) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  int x = (int)(idx % Wd);
  int y = (int)((idx / Wd) % H);
  const float* pl = in + (idx / ((long long)H * Wd)) * (long long)H * Wd;
  int ym = reflect(y - 1, H), yp = reflect(y + 1, H), xm = reflect(x - 1, Wd), xp = reflect(x + 1, Wd);
  float r0 = pl[(size_t)ym * Wd + xm] + 2.0f * pl[(size_t)ym * Wd + x] + pl[(size_t)ym * Wd + xp];
  float r1 = pl[(size_t)y * Wd + xm] + 2.0f * pl[(size_t)y * Wd + x] + pl[(size_t)y * Wd + xp];
  float r2 = pl[(size_t)yp * Wd + xm] + 2.0f * pl[(size_t)yp * Wd + x] + pl[(size_t)yp * Wd + xp];
  out[idx] = (r0 + 2.0f * r1 + r2) * 0.0625f;
}

// rgb[n][j][p] = sum_c W[j][c] net[n][c][p] + b[j] (+ prev[n][j][p]); optional sigmoid (neural_renderer.py:100,106,110).
// block = 64 pixels x 4 channel groups (the 64x64 feature map has only 4096 pixels per image: one thread per pixel left most SMs idle)
__global__ void __launch_bounds__(256) to_rgb_kernel(const float* __restrict__ W, const float* __restrict__ bias, const float* __restrict__ net,
                                                     int C, int HW, const float* __restrict__ prev, int do_sigmoid, float* __restrict__ rgb) {
  extern __shared__ float s_w[];  // [3][C] weights, then [4][64][3] partial sums
  float* s_part = s_w + 3 * C;
  for (int i = threadIdx.x; i < 3 * C; i += blockDim.x) s_w[i] = W[i];
  __syncthreads();
  const int n = blockIdx.y;
  const int px = threadIdx.x & 63, grp = threadIdx.x >> 6;
  const int p = blockIdx.x * 64 + px;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f;
  if (p < HW) {
    const float* x = net + (size_t)n * C * HW + p;
    for (int c = grp; c < C; c += 4) {
      float v = x[(size_t)c * HW];
      a0 = fmaf(s_w[c], v, a0);
      a1 = fmaf(s_w[C + c], v, a1);
      a2 = fmaf(s_w[2 * C + c], v, a2);
    }
  }
  s_part[(grp * 64 + px) * 3 + 0] = a0; s_part[(grp * 64 + px) * 3 + 1] = a1; s_part[(grp * 64 + px) * 3 + 2] = a2;
  __syncthreads();
  if (grp != 0 || p >= HW) return;
  for (int g = 1; g < 4; ++g) { a0 += s_part[(g * 64 + px) * 3 + 0]; a1 += s_part[(g * 64 + px) * 3 + 1]; a2 += s_part[(g * 64 + px) * 3 + 2]; }
  float r[3] = {a0 + bias[0], a1 + bias[1], a2 + bias[2]};
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    size_t o = ((size_t)n * 3 + j) * HW + p;
    float v = r[j];
    if (prev != nullptr) v = prev[o] + v;
    if (do_sigmoid) v = 1.0f / (1.0f + expf(-v));
    rgb[o] = v;
  }
}


// Fused all-gather of the rendered images (multi-GPU batch sharding, SURVEY §8e / BASELINE config 3): the kernel that produces the
// final RGB values also stores them into every rank's gathered buffer [3 keys][global batch][3][P*P] -- one multimem.st through the
// NVSwitch multicast address (NVLS) when available, else one st.global per peer over NVLink -- so no separate collective runs.
struct GatherDst {
  float* mc;         // multicast base of the symmetric buffer (nullptr: use peer[])
  float* peer[8];    // every rank's buffer mapped into this process (peer[rank] = own)
  int world, rank, b_local, gb;
};
__device__ __forceinline__ void multimem_st_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// net = LeakyReLU(Blur(pre)) and rgb = rgb_prev + feat_2_rgb(net) (+ sigmoid) in one pass.
// Used after an UN-activated feat conv: Blur (depthwise, weights sum to 1, reflect border) commutes with the 1x1 conv and its bias,
// so LeakyReLU(conv(Blur(x))) == LeakyReLU(Blur(conv(x)))  (neural_renderer.py:103-106, pixel_shuffle_upsample.py:7-16, 41).
// One thread = 4 consecutive x of one row, all channels (float4 loads/stores; edge taps come from the neighbours' cache lines).
// rgb_coarse (optional, instead of rgb_prev): the running RGB one level down [N][3][H/2][Wd/2]; its Blur(up2(.)) is formed here.
__global__ void __launch_bounds__(256, 3) blur_lrelu_rgb_kernel(const float* __restrict__ pre, int C, int H, int Wd,
                                                            const float* __restrict__ rgb_w, const float* __restrict__ rgb_b,
                                                            const float* __restrict__ rgb_prev, int do_sigmoid,
                                                            float* __restrict__ net, float* __restrict__ rgb, const GatherDst gd,
                                                            const float* __restrict__ rgb_coarse = nullptr) {
  extern __shared__ float s_w[];  // [3][C]
  for (int i = threadIdx.x; i < 3 * C; i += blockDim.x) s_w[i] = rgb_w[i];
  __syncthreads();
  const int n = blockIdx.y;
  const int W4 = Wd >> 2;
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= H * W4) return;
  const int y = q / W4, x0 = (q - y * W4) * 4;
  const int ym = reflect(y - 1, H), yp = reflect(y + 1, H);
  const int xl = reflect(x0 - 1, Wd), xr = reflect(x0 + 4, Wd);
  const size_t HW = (size_t)H * Wd;
  const float* pl = pre + (size_t)n * C * HW;
  float* po = net + (size_t)n * C * HW;
  float acc[3][4];
#pragma unroll
  for (int j = 0; j < 3; ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[j][i] = 0.0f;
  if (rgb_coarse != nullptr) {
    // Blur(up2(R)) of the running RGB at this thread's 4 pixels, requested up front (its scattered loads overlap the channel loop);
    // the accumulators start from it.  One set of row weights, four sets of column weights.
    int iy[3], ix[3];
    float wy[3], wx[3];
    nrf::ub_axis(y, H >> 1, iy, wy);
    const int Wc = Wd >> 1;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      nrf::ub_axis(x0 + i, Wc, ix, wx);
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const float* cp = rgb_coarse + ((size_t)n * 3 + j) * (HW >> 2);
        float v = 0.0f;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          const float* row = cp + (size_t)iy[a] * Wc;
          v += wy[a] * (wx[0] * __ldg(row + ix[0]) + wx[1] * __ldg(row + ix[1]) + wx[2] * __ldg(row + ix[2]));
        }
        acc[j][i] = v;
      }
    }
  }
  // whole-warp rows (Wd a multiple of 128): a warp holds 128 consecutive x of one row, so the outer columns x0-1 / x0+4 of a thread are
  // its neighbour lanes' inner columns (shuffles); only lanes 0 / 31 fetch theirs (one unconditional 3-row load at a per-lane column)
  const bool warp_rows = (W4 & 31) == 0;
  const int lane = threadIdx.x & 31;
  const int xe = warp_rows ? (lane == 0 ? xl : (lane == 31 ? xr : x0)) : 0;
#pragma unroll 4
  for (int c = 0; c < C; ++c) {
    const float* pc = pl + (size_t)c * HW;
    float col[6];  // vertically blurred columns x0-1 .. x0+4 (weights 1,2,1)
    {
      const float4 a = *reinterpret_cast<const float4*>(pc + (size_t)ym * Wd + x0);
      const float4 b = *reinterpret_cast<const float4*>(pc + (size_t)y * Wd + x0);
      const float4 d = *reinterpret_cast<const float4*>(pc + (size_t)yp * Wd + x0);
      col[1] = a.x + 2.0f * b.x + d.x; col[2] = a.y + 2.0f * b.y + d.y;
      col[3] = a.z + 2.0f * b.z + d.z; col[4] = a.w + 2.0f * b.w + d.w;
      if (warp_rows) {
        const float e = pc[(size_t)ym * Wd + xe] + 2.0f * pc[(size_t)y * Wd + xe] + pc[(size_t)yp * Wd + xe];
        col[0] = __shfl_up_sync(0xffffffffu, col[4], 1);
        col[5] = __shfl_down_sync(0xffffffffu, col[1], 1);
        if (lane == 0) col[0] = e;
        if (lane == 31) col[5] = e;
      } else {
        col[0] = pc[(size_t)ym * Wd + xl] + 2.0f * pc[(size_t)y * Wd + xl] + pc[(size_t)yp * Wd + xl];
        col[5] = pc[(size_t)ym * Wd + xr] + 2.0f * pc[(size_t)y * Wd + xr] + pc[(size_t)yp * Wd + xr];
      }
    }
    float o[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) o[i] = lrelu((col[i] + 2.0f * col[i + 1] + col[i + 2]) * 0.0625f);
    if (net != nullptr) *reinterpret_cast<float4*>(po + (size_t)c * HW + (size_t)y * Wd + x0) = make_float4(o[0], o[1], o[2], o[3]);
    const float w0 = s_w[c], w1 = s_w[C + c], w2 = s_w[2 * C + c];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      acc[0][i] = fmaf(w0, o[i], acc[0][i]);
      acc[1][i] = fmaf(w1, o[i], acc[1][i]);
      acc[2][i] = fmaf(w2, o[i], acc[2][i]);
    }
  }
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const size_t o = ((size_t)n * 3 + j) * HW + (size_t)y * Wd + x0;
    float v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = acc[j][i] + rgb_b[j];
    if (rgb_prev != nullptr) {
      const float4 pr = *reinterpret_cast<const float4*>(rgb_prev + o);
      v[0] += pr.x; v[1] += pr.y; v[2] += pr.z; v[3] += pr.w;
    }
    if (do_sigmoid) {
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] = 1.0f / (1.0f + expf(-v[i]));
    }
    *reinterpret_cast<float4*>(rgb + o) = make_float4(v[0], v[1], v[2], v[3]);
    if (gd.world > 1 && n < 3 * gd.b_local) {
      // local image n = key * b_local + face  ->  gathered [key][rank * b_local + face][3][HW]
      const int key = n / gd.b_local, face = n - key * gd.b_local;
      const size_t og = (((size_t)key * gd.gb + (size_t)gd.rank * gd.b_local + face) * 3 + j) * HW + (size_t)y * Wd + x0;
      if (gd.mc != nullptr) {
        multimem_st_v4(gd.mc + og, v[0], v[1], v[2], v[3]);
      } else {
        for (int pr = 0; pr < gd.world; ++pr) *reinterpret_cast<float4*>(gd.peer[pr] + og) = make_float4(v[0], v[1], v[2], v[3]);
      }
    }
  }
}

// out = Blur(bilinear_up2(in)), in [planes,H,W] -> out [planes,2H,2W]  (neural_renderer.py:65-67).
__device__ __forceinline__ float up2_at(const float* pl, int H, int Wd, int Y, int X) {
  // bilinear, align_corners=False, scale 2: src = (dst + .5)/2 - .5, clamped at 0 (PyTorch semantics)
  int y0 = (Y >> 1) - ((Y & 1) ? 0 : 1), x0 = (X >> 1) - ((X & 1) ? 0 : 1);
  float wy1 = (Y & 1) ? 0.25f : 0.75f, wx1 = (X & 1) ? 0.25f : 0.75f;  // weight of the upper neighbour (y0+1, x0+1)
  int y1 = min(y0 + 1, H - 1), x1 = min(x0 + 1, Wd - 1);
  if (y0 < 0) { y0 = 0; }
  if (x0 < 0) { x0 = 0; }
  float wy0 = 1.0f - wy1, wx0 = 1.0f - wx1;
  float top = wx0 * pl[(size_t)y0 * Wd + x0] + wx1 * pl[(size_t)y0 * Wd + x1];
  float bot = wx0 * pl[(size_t)y1 * Wd + x0] + wx1 * pl[(size_t)y1 * Wd + x1];
  return wy0 * top + wy1 * bot;
}

__global__ void rgb_up_blur_kernel(const float* __restrict__ in, int H, int Wd, long long total, float* __restrict__ out) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int H2 = 2 * H, W2 = 2 * Wd;
  int X = (int)(idx % W2);
  int Y = (int)((idx / W2) % H2);
  const float* pl = in + (idx / ((long long)H2 * W2)) * (long long)H * Wd;
  const float k[3] = {1.0f, 2.0f, 1.0f};
  float acc = 0.0f;
#pragma unroll
  for (int dy = -1; dy <= 1; ++dy) {
    int yy = reflect(Y + dy, H2);
    float row = 0.0f;
#pragma unroll
    for (int dx = -1; dx <= 1; ++dx) row += k[dx + 1] * up2_at(pl, H, Wd, yy, reflect(X + dx, W2));
    acc += k[dy + 1] * row;
  }
  out[idx] = acc * 0.0625f;
}

struct NrPlan {
  size_t t1, sh, bl, net, rgb_a, rgb_b, total;  // byte offsets / total
};

static NrPlan nr_plan(int N, int C, int S, int n_blocks, int min_feat) {
  size_t m_t1 = 0, m_sh = 0, m_net = 0;
  for (int i = 0; i < n_blocks; ++i) {
    size_t ci = (size_t)((C >> i) > min_feat ? (C >> i) : min_feat);
    size_t co = (size_t)((C >> (i + 1)) > min_feat ? (C >> (i + 1)) : min_feat);
    size_t s = (size_t)S << i;
    m_t1 = max(m_t1, (size_t)N * 2 * ci * s * s);
    m_sh = max(m_sh, (size_t)N * ci * 4 * s * s);
    m_net = max(m_net, (size_t)N * co * 4 * s * s);
  }
  size_t P = (size_t)S << n_blocks;
  size_t m_rgb = (size_t)N * 3 * P * P;
  auto al = [](size_t v) { return (v * sizeof(float) + 255) & ~(size_t)255; };
  NrPlan p;
  p.t1 = 0;
  p.sh = p.t1 + al(m_t1);
  p.bl = p.sh + al(m_sh);
  p.net = p.bl + al(m_sh);
  p.rgb_a = p.net + 2 * al(m_net);  // two net buffers (ping-pong between levels)
  p.rgb_b = p.rgb_a + al(m_rgb);
  p.total = p.rgb_b + al(m_rgb);
  return p;
}

}  // namespace gnrf

using namespace gnrf;

// ---- fused path (nr_fused.cuh): two kernels per level + the final blur / to-RGB pass ------------------------------------------
static inline int nr_width(int C, int i, int min_feat) { return (C >> i) > min_feat ? (C >> i) : min_feat; }

// the fused kernels cover this configuration: every level fits their smem / TMEM budgets and every level's pixel count is a
// multiple of the 128-pixel tile
static bool nr_fused_supported(int C, int S, int n_blocks, int min_feat) {
  if (n_blocks < 1 || n_blocks > 4 || (S * S) % nrf::kTile != 0 || S % 4 != 0) return false;
  for (int i = 0; i < n_blocks; ++i)
    if (!nrf::level_supported(nrf::level_geom(nr_width(C, i, min_feat), nr_width(C, i + 1, min_feat)))) return false;
  return true;
}
static size_t nr_fused_pack_bytes(int C, int n_blocks, int min_feat) {
  size_t tot = 0;
  for (int i = 0; i < n_blocks; ++i) tot += nrf::level_pack(nrf::level_geom(nr_width(C, i, min_feat), nr_width(C, i + 1, min_feat))).total;
  return tot;
}
struct NrFusedPlan {
  size_t t1, x, pre[2], rgb[2], rgb_up, total;
};
static NrFusedPlan nr_fused_plan(int N, int C, int S, int n_blocks, int min_feat) {
  size_t m_t1 = 0, m_x = 0, m_pre = 0;
  for (int i = 0; i < n_blocks; ++i) {
    const nrf::LevelGeom g = nrf::level_geom(nr_width(C, i, min_feat), nr_width(C, i + 1, min_feat));
    const size_t hw = ((size_t)S << i) * ((size_t)S << i);
    m_t1 = max(m_t1, (size_t)N * (hw / nrf::kTile) * g.k2_steps * nrf::kStepBytes);
    if (i > 0) m_x = max(m_x, (size_t)N * g.ci * hw * sizeof(float));
    m_pre = max(m_pre, (size_t)N * g.co * 4 * hw * sizeof(float));
  }
  const size_t P = (size_t)S << n_blocks;
  const size_t m_rgb = (size_t)N * 3 * P * P * sizeof(float);
  auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
  NrFusedPlan p;
  size_t off = 0;
  p.t1 = off; off += al(m_t1);
  p.x = off; off += al(m_x);
  p.pre[0] = off; off += al(m_pre);
  p.pre[1] = off; off += al(m_pre);
  p.rgb[0] = off; off += al(m_rgb);
  p.rgb[1] = off; off += al(m_rgb);
  p.rgb_up = off; off += al(m_rgb);
  p.total = off;
  return p;
}

extern "C" size_t gnrf_nr_workspace_bytes(int N, int C, int S, int n_blocks, int min_feat) {
  if (N <= 0 || C <= 0 || S <= 0 || n_blocks < 0) return 0;
  size_t need = nr_plan(N, C, S, n_blocks, min_feat).total;
  if (nr_fused_supported(C, S, n_blocks, min_feat)) need = max(need, nr_fused_plan(N, C, S, n_blocks, min_feat).total);
  return need;
}

// per-level buffers of one forward: aliased across levels for inference (nr_plan), distinct for training (nr_train_plan)
struct NrBufs {
  float* t1[8]; float* sh[8]; float* net[8];
  float* bl; float* rgb_a; float* rgb_b;
  bool keep_last;   // write the last level's activation (needed by the backward only)
  GatherDst gather; // world <= 1: no fused all-gather
};

static int nr_forward_bufs(const float* const* params, const unsigned char* packed, const float* featmap, int N, int C, int S, int n_blocks,
                           int min_feat, float* img, const NrBufs& bufs, cudaStream_t st);

// packed == nullptr: fp32 CUDA-core GEMMs; else: tcgen05 bf16x3 GEMMs on the packed weight streams (3 layers per block)
static size_t nr_conv_pack_bytes(int C, int n_blocks, int min_feat);

// Fused forward: per level nrf_a (x on load -> W1 GEMM -> t1 tiles + running RGB) and nrf_b (W2 GEMM -> drain -> W3 GEMM -> pre), then the
// final pass  img = sigmoid(Blur(up2(R)) + toRGB(LReLU(Blur(pre))))  (models/neural_renderer.py:98-113).
static int nr_forward_fused(const float* const* params, const unsigned char* fpack, const float* featmap, int N, int C, int S, int n_blocks,
                            int min_feat, float* img, void* workspace, size_t workspace_bytes, cudaStream_t st, const GatherDst* gather) {
  const NrFusedPlan pl = nr_fused_plan(N, C, S, n_blocks, min_feat);
  if (workspace_bytes < pl.total)
    return fail(GNRF_ERR_ARG, "gnrf_neural_render_tc_fwd: workspace %zu < required %zu bytes", workspace_bytes, pl.total);
  for (int i = 0; i < 8 * n_blocks + 2; ++i) GNRF_CHECK_ARG(params[i] != nullptr);
  int n_sm = 0;
  {
    int rc = device_once(kOnceNrFused, &n_sm, []() -> int {
      GNRF_CUDA(cudaFuncSetAttribute(nrf::nrf_a_kernel<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, nrf::kSmemBudget));
      GNRF_CUDA(cudaFuncSetAttribute(nrf::nrf_a_kernel<258, 129>, cudaFuncAttributeMaxDynamicSharedMemorySize, nrf::kSmemBudget));
      GNRF_CUDA(cudaFuncSetAttribute(nrf::nrf_a_kernel<129, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, nrf::kSmemBudget));
      GNRF_CUDA(cudaFuncSetAttribute(nrf::nrf_a_kernel<64, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, nrf::kSmemBudget));
      GNRF_CUDA(cudaFuncSetAttribute(nrf::nrf_b_kernel<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, nrf::kSmemBudget));
      GNRF_CUDA(cudaFuncSetAttribute(nrf::nrf_b_kernel<258, 129>, cudaFuncAttributeMaxDynamicSharedMemorySize, nrf::kSmemBudget));
      GNRF_CUDA(cudaFuncSetAttribute(nrf::nrf_b_kernel<129, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, nrf::kSmemBudget));
      GNRF_CUDA(cudaFuncSetAttribute(nrf::nrf_b_kernel<64, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, nrf::kSmemBudget));
      return GNRF_OK;
    });
    if (rc != GNRF_OK) return rc;
  }
  char* ws = static_cast<char*>(workspace);
  unsigned char* t1 = reinterpret_cast<unsigned char*>(ws + pl.t1);
  float* xbuf = reinterpret_cast<float*>(ws + pl.x);
  float* pre[2] = {reinterpret_cast<float*>(ws + pl.pre[0]), reinterpret_cast<float*>(ws + pl.pre[1])};
  float* rgb[2] = {reinterpret_cast<float*>(ws + pl.rgb[0]), reinterpret_cast<float*>(ws + pl.rgb[1])};
  int s = S;
  for (int i = 0; i < n_blocks; ++i) {
    const int ci = nr_width(C, i, min_feat), co = nr_width(C, i + 1, min_feat);
    const nrf::LevelGeom g = nrf::level_geom(ci, co);
    const int n_tiles = N * (s * s / nrf::kTile);
    nrf::AArgs a;
    a.src = i == 0 ? featmap : pre[(i - 1) & 1];
    a.blur = i == 0 ? 0 : 1;
    a.x_out = i == 0 ? nullptr : xbuf;
    a.pack = fpack;
    a.t1 = t1;
    a.rgb_out = rgb[i & 1];
    a.rgb_prev = i == 0 ? nullptr : rgb[(i - 1) & 1];
    a.ci = ci; a.co = co; a.H = s; a.W = s; a.n_img = N;
    // >= 120 KB of dynamic smem per CTA: one CTA per SM, so its 512-column TMEM allocation never waits for a neighbour
    const int ga = min(n_tiles, n_sm), sa = max(nrf::nrf_a_smem_bytes(g), 120 * 1024);
    if (ci == 258 && co == 129) nrf::nrf_a_kernel<258, 129><<<ga, nrf::kAThreads, sa, st>>>(a);
    else if (ci == 129 && co == 64) nrf::nrf_a_kernel<129, 64><<<ga, nrf::kAThreads, sa, st>>>(a);
    else if (ci == 64 && co == 32) nrf::nrf_a_kernel<64, 32><<<ga, nrf::kAThreads, sa, st>>>(a);
    else nrf::nrf_a_kernel<0, 0><<<ga, nrf::kAThreads, sa, st>>>(a);
    nrf::BArgs b;
    b.t1 = t1;
    b.pack = fpack;
    b.xres = i == 0 ? featmap : xbuf;
    b.pre = pre[i & 1];
    b.ci = ci; b.co = co; b.H = s; b.W = s; b.n_img = N;
    const int n_items = n_tiles * (4 / g.qg);
    const int gb = min(n_items, n_sm), sb = max(nrf::nrf_b_smem_bytes(g), 120 * 1024);
    if (ci == 258 && co == 129) nrf::nrf_b_kernel<258, 129><<<gb, nrf::kBThreads, sb, st>>>(b);
    else if (ci == 129 && co == 64) nrf::nrf_b_kernel<129, 64><<<gb, nrf::kBThreads, sb, st>>>(b);
    else if (ci == 64 && co == 32) nrf::nrf_b_kernel<64, 32><<<gb, nrf::kBThreads, sb, st>>>(b);
    else nrf::nrf_b_kernel<0, 0><<<gb, nrf::kBThreads, sb, st>>>(b);
    GNRF_LAUNCH_CHECK();
    count_launches(2);
    fpack += nrf::level_pack(g).total;
    s *= 2;
  }
  // final: Blur(up2(R_{nb-1})) (3 channels), then img = sigmoid(. + toRGB_nb(LReLU(Blur(pre_{nb-1})))), all-gather fused when requested
  {
    const int co = nr_width(C, n_blocks, min_feat);
    dim3 grid(ceil_div(s * s / 4, 256), N);
    blur_lrelu_rgb_kernel<<<grid, 256, 3 * co * sizeof(float), st>>>(pre[(n_blocks - 1) & 1], co, s, s, params[4 * n_blocks + 2 * n_blocks],
                                                                    params[4 * n_blocks + 2 * n_blocks + 1], nullptr, 1, nullptr, img,
                                                                    gather ? *gather : GatherDst{}, rgb[(n_blocks - 1) & 1]);
    GNRF_LAUNCH_CHECK();
    count_launches(1);
  }
  return GNRF_OK;
}

// allow_fused == false: the layer-wise conv_tc path (also taken for feature maps the fused kernels do not cover, i.e. fewer than 128
// pixels per image), kept as an on-device cross-check of the fused kernels
static int nr_forward(const float* const* params, const unsigned char* packed, const float* featmap, int N, int C, int S, int n_blocks,
                      int min_feat, float* img, void* workspace, size_t workspace_bytes, gnrf_stream_t stream,
                      const GatherDst* gather = nullptr, bool allow_fused = true) {
  GNRF_CHECK_ARG(params && featmap && img && workspace);
  GNRF_CHECK_ARG(N > 0 && C > 0 && S >= 2 && n_blocks >= 1 && n_blocks <= 6);
  GNRF_CHECK_ARG(S % 4 == 0);
  if (packed != nullptr && allow_fused && nr_fused_supported(C, S, n_blocks, min_feat))
    return nr_forward_fused(params, packed + nr_conv_pack_bytes(C, n_blocks, min_feat), featmap, N, C, S, n_blocks, min_feat, img, workspace,
                            workspace_bytes, as_stream(stream), gather);
  NrPlan pl = nr_plan(N, C, S, n_blocks, min_feat);
  if (workspace_bytes < pl.total)
    return fail(GNRF_ERR_ARG, "gnrf_neural_render_fwd: workspace %zu < required %zu bytes", workspace_bytes, pl.total);
  char* ws = static_cast<char*>(workspace);
  NrBufs bufs;
  float* netbuf[2] = {reinterpret_cast<float*>(ws + pl.net), reinterpret_cast<float*>(ws + pl.net + (pl.rgb_a - pl.net) / 2)};
  for (int i = 0; i < n_blocks; ++i) {
    bufs.t1[i] = reinterpret_cast<float*>(ws + pl.t1);
    bufs.sh[i] = reinterpret_cast<float*>(ws + pl.sh);
    bufs.net[i] = netbuf[i & 1];
  }
  bufs.bl = reinterpret_cast<float*>(ws + pl.bl);
  bufs.rgb_a = reinterpret_cast<float*>(ws + pl.rgb_a);
  bufs.rgb_b = reinterpret_cast<float*>(ws + pl.rgb_b);
  bufs.keep_last = false;
  bufs.gather = gather ? *gather : GatherDst{};
  return nr_forward_bufs(params, packed, featmap, N, C, S, n_blocks, min_feat, img, bufs, as_stream(stream));
}

static int nr_forward_bufs(const float* const* params, const unsigned char* packed, const float* featmap, int N, int C, int S, int n_blocks,
                           int min_feat, float* img, const NrBufs& bufs, cudaStream_t st) {
  float* bl = bufs.bl;
  float* rgb_a = bufs.rgb_a;
  float* rgb_b = bufs.rgb_b;

  // parameter indexing (see gnrf.h): psu i: [4i..4i+3]; to_rgb j: [4nb + 2j, +1]; feat i: [4nb + 2(nb+1) + 2i, +1]
  auto psu_w = [&](int i, int l) { return params[4 * i + 2 * l]; };
  auto psu_b = [&](int i, int l) { return params[4 * i + 2 * l + 1]; };
  auto rgb_w = [&](int j) { return params[4 * n_blocks + 2 * j]; };
  auto rgb_bi = [&](int j) { return params[4 * n_blocks + 2 * j + 1]; };
  auto feat_w = [&](int i) { return params[4 * n_blocks + 2 * (n_blocks + 1) + 2 * i]; };
  auto feat_b = [&](int i) { return params[4 * n_blocks + 2 * (n_blocks + 1) + 2 * i + 1]; };
  for (int i = 0; i < 8 * n_blocks + 2; ++i) GNRF_CHECK_ARG(params[i] != nullptr);

  auto launch_rgb = [&](const float* w, const float* b, const float* net, int Cn, int HW, const float* prev, int sig, float* dst) {
    dim3 grid(ceil_div(HW, 64), N);
    to_rgb_kernel<<<grid, 256, (3 * Cn + 4 * 64 * 3) * sizeof(float), st>>>(w, b, net, Cn, HW, prev, sig, dst);
  };
  auto launch_up = [&](const float* src, int H, float* dst) {
    long long total = (long long)N * 3 * 4 * H * H;
    rgb_up_blur_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(src, H, H, total, dst);
  };

  int s = S;
  // rgb = rgb_upsample(feat_2_rgb_list[0](x))
  launch_rgb(rgb_w(0), rgb_bi(0), featmap, C, s * s, nullptr, 0, rgb_a);
  launch_up(rgb_a, s, rgb_b);
  GNRF_LAUNCH_CHECK();
  count_launches(2);
  const float* net = featmap;
  size_t pk_off = 0;
  float* rgb_up = rgb_b;   // upsampled running rgb
  float* rgb_cur = rgb_a;  // scratch for the next sum
  for (int i = 0; i < n_blocks; ++i) {
    int ci = (C >> i) > min_feat ? (C >> i) : min_feat;
    int co = (C >> (i + 1)) > min_feat ? (C >> (i + 1)) : min_feat;
    int HW = s * s;
    tc::ConvLayerPlan pl1 = tc::conv_layer_plan(2 * ci, ci), pl2 = tc::conv_layer_plan(4 * ci, 2 * ci), pl3 = tc::conv_layer_plan(co, ci);
    float* t1 = bufs.t1[i];
    float* sh = bufs.sh[i];
    if (packed != nullptr) {
      int rc = tc::conv_tc_launch(pl1, packed + pk_off, net, t1, nullptr, 1, N, HW, s, tc::CONV_EPI_LRELU, st);
      if (rc != GNRF_OK) return rc;
      rc = tc::conv_tc_launch(pl2, packed + pk_off + pl1.total_bytes, t1, sh, net, ci, N, HW, s, tc::CONV_EPI_PSU, st);
      if (rc != GNRF_OK) return rc;
    } else {
      {  // PSU layer_1: ci -> 2ci, LeakyReLU
        dim3 grid(ceil_div(HW, BN), ceil_div(2 * ci, BM), N);
        conv1x1_kernel<EPI_LRELU><<<grid, kGemmThreads, 0, st>>>(psu_w(i, 0), psu_b(i, 0), net, 2 * ci, ci, HW, s, nullptr, 1, t1);
      }
      {  // PSU layer_2: 2ci -> 4ci, LeakyReLU, + repeat(x,4), pixel_shuffle(2)
        dim3 grid(ceil_div(HW, BN), ceil_div(4 * ci, BM), N);
        conv1x1_kernel<EPI_PSU><<<grid, kGemmThreads, 0, st>>>(psu_w(i, 1), psu_b(i, 1), t1, 4 * ci, 2 * ci, HW, s, net, ci, sh);
      }
      count_launches(2);
    }
    s *= 2;
    HW = s * s;
    float* net_out = bufs.net[i];
    bool last = (i == n_blocks - 1);
    if (packed != nullptr) {
      // feat_layers[i] un-activated on tensor cores (into `bl`), then ONE pass: net = LeakyReLU(Blur(.)), rgb += feat_2_rgb(net)
      int rc = tc::conv_tc_launch(pl3, packed + pk_off + pl1.total_bytes + pl2.total_bytes, sh, bl, nullptr, 1, N, HW, s,
                                  tc::CONV_EPI_LINEAR, st);
      if (rc != GNRF_OK) return rc;
      dim3 grid(ceil_div(HW / 4, 256), N);
      // the last level's activation is consumed only by the fused to-RGB head: not written unless the caller keeps it (training)
      blur_lrelu_rgb_kernel<<<grid, 256, 3 * co * sizeof(float), st>>>(bl, co, s, s, rgb_w(i + 1), rgb_bi(i + 1), rgb_up, last ? 1 : 0,
                                                                      (last && !bufs.keep_last) ? nullptr : net_out, last ? img : rgb_cur,
                                                                      last ? bufs.gather : GatherDst{});
      count_launches(1);
    } else {
      {
        long long total = (long long)N * ci * HW;
        blur3x3_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(sh, s, s, total, bl);
      }
      {  // feat_layers[i]: ci -> co, LeakyReLU
        dim3 grid(ceil_div(HW, BN), ceil_div(co, BM), N);
        conv1x1_kernel<EPI_LRELU><<<grid, kGemmThreads, 0, st>>>(feat_w(i), feat_b(i), bl, co, ci, HW, s, nullptr, 1, net_out);
      }
      launch_rgb(rgb_w(i + 1), rgb_bi(i + 1), net_out, co, HW, rgb_up, last ? 1 : 0, last ? img : rgb_cur);
      count_launches(3);
    }
    pk_off += pl1.total_bytes + pl2.total_bytes + pl3.total_bytes;
    if (!last) {
      float* dst = (rgb_cur == rgb_a) ? rgb_b : rgb_a;  // == rgb_up's buffer, whose contents are now consumed
      launch_up(rgb_cur, s, dst);
      rgb_up = dst;
      rgb_cur = (dst == rgb_a) ? rgb_b : rgb_a;
      count_launches(1);
    }
    net = net_out;
    GNRF_LAUNCH_CHECK();
  }
  return GNRF_OK;
}

extern "C" int gnrf_neural_render_fwd(const float* const* params, const float* featmap, int N, int C, int S, int n_blocks,
                                      int min_feat, float* img, void* workspace, size_t workspace_bytes, gnrf_stream_t stream) {
  return nr_forward(params, nullptr, featmap, N, C, S, n_blocks, min_feat, img, workspace, workspace_bytes, stream);
}

// packed = [per-layer conv_tc streams (layer-wise path: training forward, small feature maps)] [fused per-level images (nr_fused.cuh)]
static size_t gnrf_nr_tc_packed_bytes_conv(int C, int n_blocks, int min_feat) {
  size_t tot = 0;
  for (int i = 0; i < n_blocks; ++i) {
    int ci = (C >> i) > min_feat ? (C >> i) : min_feat;
    int co = (C >> (i + 1)) > min_feat ? (C >> (i + 1)) : min_feat;
    tot += tc::conv_layer_plan(2 * ci, ci).total_bytes + tc::conv_layer_plan(4 * ci, 2 * ci).total_bytes + tc::conv_layer_plan(co, ci).total_bytes;
  }
  return tot;
}
static size_t nr_conv_pack_bytes(int C, int n_blocks, int min_feat) { return gnrf_nr_tc_packed_bytes_conv(C, n_blocks, min_feat); }

extern "C" size_t gnrf_nr_tc_packed_bytes(int C, int n_blocks, int min_feat) {
  if (C <= 0 || n_blocks < 1 || n_blocks > 6) return 0;
  return nr_conv_pack_bytes(C, n_blocks, min_feat) + nr_fused_pack_bytes(C, n_blocks, min_feat);
}

extern "C" int gnrf_nr_tc_pack(const float* const* params, int C, int n_blocks, int min_feat, void* packed, gnrf_stream_t stream) {
  GNRF_CHECK_ARG(params && packed && C > 0 && n_blocks >= 1 && n_blocks <= 6);
  GNRF_CHECK_ARG((reinterpret_cast<uintptr_t>(packed) & 127) == 0);
  unsigned char* dst = static_cast<unsigned char*>(packed);
  for (int i = 0; i < n_blocks; ++i) {
    int ci = (C >> i) > min_feat ? (C >> i) : min_feat;
    int co = (C >> (i + 1)) > min_feat ? (C >> (i + 1)) : min_feat;
    tc::ConvLayerPlan pl[3] = {tc::conv_layer_plan(2 * ci, ci), tc::conv_layer_plan(4 * ci, 2 * ci), tc::conv_layer_plan(co, ci)};
    const float* w[3] = {params[4 * i], params[4 * i + 2], params[4 * n_blocks + 2 * (n_blocks + 1) + 2 * i]};
    const float* b[3] = {params[4 * i + 1], params[4 * i + 3], params[4 * n_blocks + 2 * (n_blocks + 1) + 2 * i + 1]};
    for (int l = 0; l < 3; ++l) {
      GNRF_CHECK_ARG(w[l] && b[l]);
      int rc = tc::conv_tc_pack(pl[l], w[l], b[l], dst, as_stream(stream));
      if (rc != GNRF_OK) return rc;
      dst += pl[l].total_bytes;
    }
  }
  // fused per-level images (weights of level i: PSU layer_1 / layer_2, feat_layers[i], and the to-RGB head of the level's INPUT)
  for (int i = 0; i < n_blocks; ++i) {
    int ci = (C >> i) > min_feat ? (C >> i) : min_feat;
    int co = (C >> (i + 1)) > min_feat ? (C >> (i + 1)) : min_feat;
    nrf::PackArgs pa;
    pa.w1 = params[4 * i]; pa.b1 = params[4 * i + 1]; pa.w2 = params[4 * i + 2]; pa.b2 = params[4 * i + 3];
    pa.w3 = params[4 * n_blocks + 2 * (n_blocks + 1) + 2 * i]; pa.b3 = params[4 * n_blocks + 2 * (n_blocks + 1) + 2 * i + 1];
    pa.wrgb = params[4 * n_blocks + 2 * i]; pa.brgb = params[4 * n_blocks + 2 * i + 1];
    GNRF_CHECK_ARG(pa.wrgb && pa.brgb);
    pa.dst = dst;
    nrf::nrf_pack_kernel<<<148, 256, 0, as_stream(stream)>>>(pa, ci, co);
    GNRF_LAUNCH_CHECK();
    count_launches(1);
    dst += nrf::level_pack(nrf::level_geom(ci, co)).total;
  }
  return GNRF_OK;
}

extern "C" int gnrf_neural_render_tc_fwd(const float* const* params, const void* packed, const float* featmap, int N, int C, int S,
                                         int n_blocks, int min_feat, float* img, void* workspace, size_t workspace_bytes,
                                         gnrf_stream_t stream) {
  GNRF_CHECK_ARG(packed);
  return nr_forward(params, static_cast<const unsigned char*>(packed), featmap, N, C, S, n_blocks, min_feat, img, workspace,
                    workspace_bytes, stream);
}

extern "C" int gnrf_neural_render_tc_layerwise_fwd(const float* const* params, const void* packed, const float* featmap, int N, int C,
                                                   int S, int n_blocks, int min_feat, float* img, void* workspace, size_t workspace_bytes,
                                                   gnrf_stream_t stream) {
  GNRF_CHECK_ARG(packed);
  return nr_forward(params, static_cast<const unsigned char*>(packed), featmap, N, C, S, n_blocks, min_feat, img, workspace,
                    workspace_bytes, stream, nullptr, false);
}

// ------------------------------------------------------------------------------------------------- training forward
namespace gnrf {
NrTrainPlan nr_train_plan(int N, int C, int S, int n_blocks, int min_feat) {
  NrTrainPlan p;
  auto al = [](size_t v) { return (v * sizeof(float) + 255) & ~(size_t)255; };
  size_t off = 0, m_bl = 0;
  for (int i = 0; i < n_blocks; ++i) {
    size_t ci = (size_t)((C >> i) > min_feat ? (C >> i) : min_feat);
    size_t co = (size_t)((C >> (i + 1)) > min_feat ? (C >> (i + 1)) : min_feat);
    size_t s = (size_t)S << i;
    p.t1[i] = off; off += al((size_t)N * 2 * ci * s * s);
    p.sh[i] = off; off += al((size_t)N * ci * 4 * s * s);
    p.net[i] = off; off += al((size_t)N * co * 4 * s * s);
    m_bl = max(m_bl, (size_t)N * co * 4 * s * s);
  }
  size_t P = (size_t)S << n_blocks;
  p.bl = off; off += al(m_bl);
  p.rgb_a = off; off += al((size_t)N * 3 * P * P);
  p.rgb_b = off; off += al((size_t)N * 3 * P * P);
  p.total = off;
  return p;
}
}  // namespace gnrf

extern "C" size_t gnrf_nr_train_saved_bytes(int N, int C, int S, int n_blocks, int min_feat) {
  if (N <= 0 || C <= 0 || S <= 0 || n_blocks < 1 || n_blocks > 6) return 0;
  return nr_train_plan(N, C, S, n_blocks, min_feat).total;
}

// Same computation as gnrf_neural_render_tc_fwd, but every level keeps its own t1 / sh / net buffers inside `saved` so that
// gnrf_nr_train_bwd can differentiate it (layout: nr_train_plan).
extern "C" int gnrf_nr_train_fwd(const float* const* params, const void* packed, const float* featmap, int N, int C, int S, int n_blocks,
                                 int min_feat, float* img, void* saved, size_t saved_bytes, gnrf_stream_t stream) {
  GNRF_CHECK_ARG(params && packed && featmap && img && saved);
  GNRF_CHECK_ARG(N > 0 && C > 0 && S >= 2 && n_blocks >= 1 && n_blocks <= 6 && S % 4 == 0);
  NrTrainPlan pl = nr_train_plan(N, C, S, n_blocks, min_feat);
  if (saved_bytes < pl.total) return fail(GNRF_ERR_ARG, "gnrf_nr_train_fwd: saved %zu < required %zu bytes", saved_bytes, pl.total);
  char* ws = static_cast<char*>(saved);
  NrBufs bufs;
  for (int i = 0; i < n_blocks; ++i) {
    bufs.t1[i] = reinterpret_cast<float*>(ws + pl.t1[i]);
    bufs.sh[i] = reinterpret_cast<float*>(ws + pl.sh[i]);
    bufs.net[i] = reinterpret_cast<float*>(ws + pl.net[i]);
  }
  bufs.bl = reinterpret_cast<float*>(ws + pl.bl);
  bufs.rgb_a = reinterpret_cast<float*>(ws + pl.rgb_a);
  bufs.rgb_b = reinterpret_cast<float*>(ws + pl.rgb_b);
  bufs.keep_last = true;
  bufs.gather = GatherDst{};
  return nr_forward_bufs(params, static_cast<const unsigned char*>(packed), featmap, N, C, S, n_blocks, min_feat, img, bufs,
                         as_stream(stream));
}

// Same as gnrf_neural_render_tc_fwd, plus the fused all-gather of the first 3 * b_local images (see GatherDst): peer_ptrs = HOST array of
// `world` device pointers (every rank's symmetric buffer [3][gb][3][P*P] mapped into this process), mc_ptr = its multicast address or NULL.
// The caller orders buffer reuse / consumption across ranks (gazenerf_b200/dist.py: double buffering + a device barrier per step).
extern "C" int gnrf_neural_render_tc_fwd_gather(const float* const* params, const void* packed, const float* featmap, int N, int C, int S,
                                                int n_blocks, int min_feat, float* img, void* workspace, size_t workspace_bytes,
                                                float* const* peer_ptrs, float* mc_ptr, int world, int rank, int b_local, int gb,
                                                gnrf_stream_t stream) {
  GNRF_CHECK_ARG(packed && peer_ptrs);
  GNRF_CHECK_ARG(world >= 2 && world <= 8 && rank >= 0 && rank < world && b_local >= 1 && gb == world * b_local && N >= 3 * b_local);
  GatherDst gd;
  gd.mc = mc_ptr;
  for (int i = 0; i < 8; ++i) gd.peer[i] = i < world ? peer_ptrs[i] : nullptr;
  for (int i = 0; i < world; ++i) GNRF_CHECK_ARG(gd.peer[i] != nullptr);
  gd.world = world; gd.rank = rank; gd.b_local = b_local; gd.gb = gb;
  return nr_forward(params, static_cast<const unsigned char*>(packed), featmap, N, C, S, n_blocks, min_feat, img, workspace,
                    workspace_bytes, stream, &gd);
}
