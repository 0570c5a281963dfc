// Fused PixelShuffleUpsample levels of the 2-D neural renderer for sm_100a (tcgen05 / TMEM / TMA bulk copies).
// Included by neural_render.cu (one translation unit owns every neural-renderer kernel).
//
// Reference graph of one level i (models/neural_renderer.py:103-106, models/pixel_shuffle_upsample.py:33-42), x = net_i [ci, s, s]:
//     t1  = LReLU(W1 x + b1)                       ci -> 2ci          per pixel
//     t2  = LReLU(W2 t1 + b2) + repeat(x, 4)       2ci -> 4ci         per pixel
//     sh  = pixel_shuffle(t2, 2)                   [ci, 2s, 2s]       sh[c][2h+qy][2w+qx] = t2[4c + 2qy + qx][h][w]
//     net_{i+1} = LReLU(W3 Blur(sh) + b3)          ci -> co           (Blur = depthwise 3x3 binomial, reflect border)
//     rgb += toRGB_{i+1}(net_{i+1})
// Exact rewrites: Blur (depthwise, weights sum to 1) commutes with the 1x1 conv W3 and its bias, so pre_i = W3 sh + b3 is computed
// per INPUT pixel (everything up to pre_i is pixel-local) and Blur + LReLU are applied when the next stage loads pre_i; the rows of W2
// are permuted at pack time so that the 4 sub-pixels q = 2qy + qx of a pixel are contiguous column groups of t2; toRGB_i(net_i) is
// linear in the level's input and rides along as 3 extra output rows of the W1 GEMM.
//
// Two kernels per level (the 516 / 1032-wide intermediates of level 0 do not fit one CTA's smem + TMEM in split precision):
//   nrf_a_kernel : x = LReLU(Blur(pre_{i-1})) formed ON LOAD (level 0: the feature map itself) -> A tile (bf16 hi/lo) -> W1 GEMM (+ toRGB
//                  rows) -> t1 written to HBM already in UMMA operand form (hi/lo, no-swizzle K-major core matrices), running RGB R_i
//   nrf_b_kernel : t1 tiles streamed by TMA bulk copies (no conversion) -> W2 GEMM for one sub-pixel group -> drain (LReLU, + residual,
//                  hi/lo split) straight into the A operand of the W3 GEMM in smem -> pre_i stored pixel-shuffled.  t2 / sh (the widest
//                  tensors of the renderer, 67 MB per image at the last level) never exist in HBM.
// Precision: bf16x3 split (x*w ~= x_hi*w_hi + x_lo*w_hi + x_hi*w_lo, fp32 accumulate in TMEM), as everywhere in libgnrf.
#pragma once
#include "common.cuh"
#include "sm100_ptx.cuh"

namespace gnrf {
namespace nrf {

using namespace ptx;

constexpr int kTile = 128;            // pixels per tile == UMMA M
constexpr int kStepBytes = 8192;      // one K16 step of an A operand tile: [hi 4 KB | lo 4 KB], 128 rows x 32 B each
constexpr int kMaxChunks = 3;

__host__ __device__ constexpr int ceil16(int v) { return (v + 15) / 16 * 16; }

// ---- per-level geometry (ci = input channels of the level) ---------------------------------------------------------
struct LevelGeom {
  int ci, co;
  int k1_steps;                 // K16 steps of the W1 GEMM (K = ci)
  int n1, n1p;                  // 2ci + 3 (toRGB rows appended), padded to 16
  int n_chunks, chunk_n[kMaxChunks];   // N chunks of the W1 GEMM (each <= 256, multiple of 16, sum = n1p)
  int k2_steps;                 // K16 steps of t1 / the W2 GEMM (K = 2ci)
  int cip;                      // ci padded to 16 = columns of one sub-pixel group of t2 = K of the W3 GEMM
  int qg;                       // sub-pixels per nrf_b work item (4 / qg items per tile)
  int n2, n2a, n2b;             // N of the W2 GEMM per item (= qg * cip) and its split into <= 2 UMMA instructions
  int k3_steps;                 // cip / 16
  int cop;                      // co padded to 16 = N of the W3 GEMM
};

__host__ __device__ inline LevelGeom level_geom(int ci, int co) {
  LevelGeom g;
  g.ci = ci; g.co = co;
  g.k1_steps = (ci + 15) / 16;
  g.n1 = 2 * ci + 3;
  g.n1p = ceil16(g.n1);
  g.k2_steps = (2 * ci + 15) / 16;
  g.cip = ceil16(ci);
  g.k3_steps = g.cip / 16;
  g.cop = ceil16(co);
  // W1 GEMM chunks: as few as possible, each <= 256 and a multiple of 16
  g.n_chunks = (g.n1p + 255) / 256;   // two accumulator buffers at TMEM columns 0 / 256
  {
    int left = g.n1p;
    for (int i = 0; i < kMaxChunks; ++i) g.chunk_n[i] = 0;
    for (int i = 0; i < g.n_chunks; ++i) {
      int per = ceil16((left + (g.n_chunks - i) - 1) / (g.n_chunks - i));
      if (per > left) per = left;
      g.chunk_n[i] = per;
      left -= per;
    }
  }
  // sub-pixel groups per item: all four when they fit one 256-column accumulator, else one
  g.qg = (4 * g.cip <= 256) ? 4 : 1;
  g.n2 = g.qg * g.cip;
  g.n2a = g.n2 <= 256 ? g.n2 : ceil16(g.n2 / 2);
  g.n2b = g.n2 - g.n2a;
  return g;
}

// ---- packed weight image of one level ---------------------------------------------------------------------------------
//   [W1 stream][W2 stream][W3 stream][bias1 n1p][bias2 4*cip][bias3 cop]      (streams 128-byte aligned)
// W1 stream: [chunk][k16 (k1_steps)][hi | lo][chunk_n rows x 32 B]     rows: W1 (2ci), toRGB (3), zero pad
// W2 stream: [group (4/qg)][k16 (k2_steps)][hi | lo][n2 rows x 32 B]   row (q' * cip + c) of group g = W2 row 4c + (g*qg + q')
// W3 stream: [k16 (k3_steps)][hi | lo][cop rows x 32 B]
// every [rows x 32 B] slice is in the no-swizzle K-major core-matrix order: 16-byte chunk index = (row/8)*16 + k_half*8 + (row%8)
struct LevelPack {
  size_t w1, w2, w3, b1, b2, b3, total;
};
__host__ __device__ inline LevelPack level_pack(const LevelGeom& g) {
  LevelPack p;
  size_t off = 0;
  p.w1 = off; off += (size_t)g.k1_steps * 2 * g.n1p * 32;
  off = (off + 127) & ~(size_t)127;
  p.w2 = off; off += (size_t)(4 / g.qg) * g.k2_steps * 2 * g.n2 * 32;
  off = (off + 127) & ~(size_t)127;
  p.w3 = off; off += (size_t)g.k3_steps * 2 * g.cop * 32;
  off = (off + 127) & ~(size_t)127;
  p.b1 = off; off += (size_t)(g.n1p + 32) * 4;
  p.b2 = off; off += (size_t)(4 * g.cip + 32) * 4;
  p.b3 = off; off += (size_t)(g.cop + 32) * 4;
  p.total = (off + 255) & ~(size_t)255;
  return p;
}

struct PackArgs {
  const float *w1, *b1, *w2, *b2, *w3, *b3, *wrgb, *brgb;
  unsigned char* dst;
};

// one thread per 16-byte chunk of the three streams, then the biases
__global__ void nrf_pack_kernel(PackArgs a, int ci, int co) {
  const LevelGeom g = level_geom(ci, co);
  const LevelPack lp = level_pack(g);
  const size_t n_w1 = (size_t)g.k1_steps * 2 * g.n1p * 2, n_w2 = (size_t)(4 / g.qg) * g.k2_steps * 2 * g.n2 * 2,
               n_w3 = (size_t)g.k3_steps * 2 * g.cop * 2;
  const size_t total = n_w1 + n_w2 + n_w3;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    int which;
    size_t c = t;
    if (c < n_w1) which = 0;
    else if (c < n_w1 + n_w2) { which = 1; c -= n_w1; }
    else { which = 2; c -= n_w1 + n_w2; }
    int rows, k16, half, row, k_half;
    size_t dst_off;
    float vals[8];
    if (which == 0) {
      // locate the chunk: slices of chunk i have chunk_n[i] rows
      size_t rem = c;
      int chunk = 0, row0 = 0;
      for (; chunk < g.n_chunks; ++chunk) {
        const size_t in_chunk = (size_t)g.k1_steps * 2 * g.chunk_n[chunk] * 2;
        if (rem < in_chunk) break;
        rem -= in_chunk;
        row0 += g.chunk_n[chunk];
      }
      rows = g.chunk_n[chunk];
      const size_t per_slice = (size_t)rows * 2;
      const size_t s = rem / per_slice;
      const int r16 = (int)(rem % per_slice);
      half = (int)(s & 1); k16 = (int)(s >> 1);
      row = (r16 >> 4) * 8 + (r16 & 7); k_half = (r16 >> 3) & 1;
      const int n = row0 + row;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int k = k16 * 16 + k_half * 8 + j;
        float v = 0.0f;
        if (k < ci) {
          if (n < 2 * ci) v = a.w1[(size_t)n * ci + k];
          else if (n < 2 * ci + 3) v = a.wrgb[(size_t)(n - 2 * ci) * ci + k];
        }
        vals[j] = v;
      }
      dst_off = lp.w1 + (size_t)(t) * 16;   // W1 stream chunks are laid out exactly in thread order
    } else if (which == 1) {
      rows = g.n2;
      const size_t per_slice = (size_t)rows * 2;
      size_t s = c / per_slice;
      const int r16 = (int)(c % per_slice);
      half = (int)(s & 1); s >>= 1;
      k16 = (int)(s % g.k2_steps);
      const int grp = (int)(s / g.k2_steps);
      row = (r16 >> 4) * 8 + (r16 & 7); k_half = (r16 >> 3) & 1;
      const int qq = row / g.cip, cc = row - qq * g.cip;
      const int q = grp * g.qg + qq;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int k = k16 * 16 + k_half * 8 + j;
        vals[j] = (cc < ci && k < 2 * ci) ? a.w2[(size_t)(4 * cc + q) * (2 * ci) + k] : 0.0f;
      }
      dst_off = lp.w2 + c * 16;
    } else {
      rows = g.cop;
      const size_t per_slice = (size_t)rows * 2;
      const size_t s = c / per_slice;
      const int r16 = (int)(c % per_slice);
      half = (int)(s & 1); k16 = (int)(s >> 1);
      row = (r16 >> 4) * 8 + (r16 & 7); k_half = (r16 >> 3) & 1;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int k = k16 * 16 + k_half * 8 + j;
        vals[j] = (row < co && k < ci) ? a.w3[(size_t)row * ci + k] : 0.0f;
      }
      dst_off = lp.w3 + c * 16;
    }
    uint32_t out[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint32_t hi, lo;
      split2(vals[2 * j], vals[2 * j + 1], hi, lo);
      out[j] = half ? lo : hi;
    }
    *reinterpret_cast<uint4*>(a.dst + dst_off) = make_uint4(out[0], out[1], out[2], out[3]);
  }
  float* b1 = reinterpret_cast<float*>(a.dst + lp.b1);
  float* b2 = reinterpret_cast<float*>(a.dst + lp.b2);
  float* b3 = reinterpret_cast<float*>(a.dst + lp.b3);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < g.n1p + 32; i += gridDim.x * blockDim.x)
    b1[i] = i < 2 * ci ? a.b1[i] : (i < 2 * ci + 3 ? a.brgb[i - 2 * ci] : 0.0f);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 4 * g.cip + 32; i += gridDim.x * blockDim.x) {
    const int q = i / g.cip, c = i - q * g.cip;
    b2[i] = (q < 4 && c < ci) ? a.b2[4 * c + q] : 0.0f;
  }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < g.cop + 32; i += gridDim.x * blockDim.x) b3[i] = i < co ? a.b3[i] : 0.0f;
}

__device__ __forceinline__ float lrelu02(float v) { return fmaxf(v, 0.2f * v); }
__device__ __forceinline__ int reflect_idx(int i, int n) { return i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i); }

// descriptor halves for no-swizzle K-major operands (A and B): LBO = 128 B (K halves), SBO = 256 B (8-row groups), version 1
constexpr uint32_t kDescHiNoSw = (uint32_t)((256 >> 4) | (1u << 14));
constexpr uint32_t kDescLoLboNo = (128u >> 4) << 16;
__device__ __forceinline__ uint64_t mk_desc(uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; }

// =====================================================================================================================
//  nrf_a_kernel
// =====================================================================================================================
struct AArgs {
  const float* src;        // level 0: feature map [N][ci][HW]; else pre_{i-1} [N][ci][H][W] (un-blurred, un-activated)
  int blur;                // 0: x = src; 1: x = LReLU(Blur3x3(src)) (reflect border)
  float* x_out;            // blur mode: x [N][ci][HW] (the level's activated input, read again as the PSU residual); may be null
  const unsigned char* pack;
  unsigned char* t1;       // [tile][k2_steps][hi 4 KB | lo 4 KB]
  float* rgb_out;          // R_i [N][3][HW] = toRGB_i(x) + b (+ Blur(up2(R_{i-1})))
  const float* rgb_prev;   // R_{i-1} [N][3][HW/4] or null
  int ci, co, H, W, n_img;
};

constexpr int kABStages = 6;                             // ring of W1 K16 slices: [chunk_n rows x 32 B] hi + lo per stage
__host__ __device__ inline int nrf_a_bstage_bytes(const LevelGeom& g) { return (g.chunk_n[0] * 64 + 127) & ~127; }   // chunk 0 is the widest
__host__ __device__ inline int nrf_a_nbuf(const LevelGeom& g) {
  return (2 * g.k1_steps * kStepBytes + kABStages * nrf_a_bstage_bytes(g) <= 216 * 1024) ? 2 : 1;
}
constexpr int kAThreads = 448;                           // warps 0-3 epilogue, 4-11 loaders, 12 weight TMA, 13 MMA
constexpr int kABarAFull = 0, kABarAEmpty = 2, kABarBFull = 4, kABarBEmpty = kABarBFull + kABStages, kABarAccFull = kABarBEmpty + kABStages,
              kABarAccEmpty = kABarAccFull + 2, kANumBars = kABarAccEmpty + 2;

// Blur(up2(prev)) at fine pixel (Y, X): bilinear x2 (align_corners=False, clamped) followed by the 3x3 binomial with reflect border
// (models/neural_renderer.py:65-67) is, per axis, a 3-tap stencil on the COARSE grid: fine index F reads coarse indices
// {c-1, c, c+1} (c = F >> 1, clamped) with weights accumulated from the three blur taps (1,2,1)/4 at reflect(F-1), F, reflect(F+1), each
// of which is a two-tap interpolation (.25/.75) of clamped coarse neighbours.  Interior: (1.25, 2.5, .25)/4 for even F, mirrored for odd.
__device__ __forceinline__ void ub_axis(int F, int n_coarse, int (&idx)[3], float (&w)[3]) {
  const int n_fine = 2 * n_coarse;
  const int c = F >> 1;
  idx[0] = max(c - 1, 0); idx[1] = c; idx[2] = min(c + 1, n_coarse - 1);
  if (F >= 2 && F <= n_fine - 3) {   // interior: no clamp / reflect is hit
    const bool odd = F & 1;
    w[0] = odd ? 0.0625f : 0.3125f; w[1] = 0.625f; w[2] = odd ? 0.3125f : 0.0625f;
    return;
  }
  w[0] = w[1] = w[2] = 0.0f;
  const float kf[3] = {0.25f, 0.5f, 0.25f};
#pragma unroll
  for (int d = -1; d <= 1; ++d) {
    const int f = reflect_idx(F + d, n_fine);
    int i0 = (f >> 1) - ((f & 1) ? 0 : 1);
    const float w1 = (f & 1) ? 0.25f : 0.75f;
    const int i1 = min(i0 + 1, n_coarse - 1);
    i0 = max(i0, 0);
    // clamped neighbours always fall inside {c-1, c, c+1}; slot = i - c + 1 (a clamped index lands on the slot that holds it)
    const int s0 = min(max(i0 - c + 1, 0), 2), s1 = min(max(i1 - c + 1, 0), 2);
#pragma unroll
    for (int t = 0; t < 3; ++t) {
      if (t == s0) w[t] += kf[d + 1] * (1.0f - w1);
      if (t == s1) w[t] += kf[d + 1] * w1;
    }
  }
}
__device__ __forceinline__ float up2_blur_at(const float* pl, int H, int W, int Y, int X) {
  int iy[3], ix[3];
  float wy[3], wx[3];
  ub_axis(Y, H, iy, wy);
  ub_axis(X, W, ix, wx);
  float acc = 0.0f;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float* row = pl + (size_t)iy[a] * W;
    acc += wy[a] * (wx[0] * __ldg(row + ix[0]) + wx[1] * __ldg(row + ix[1]) + wx[2] * __ldg(row + ix[2]));
  }
  return acc;
}

// out = Blur(up2(in)) for [planes][H][W] -> [planes][2H][2W]; grid = (ceil(4HW / 256), planes), 32-bit index math only
__global__ void ub_kernel(const float* __restrict__ in, int H, int W, float* __restrict__ out) {
  const int W2 = 2 * W, n = 4 * H * W;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int Y = i / W2, X = i - Y * W2;
  out[(size_t)blockIdx.y * n + i] = up2_blur_at(in + (size_t)blockIdx.y * H * W, H, W, Y, X);
}

// CI > 0: the level's channel counts are compile-time constants (the reference configuration 258 / 129 / 64): every division and
// modulo of the index math and the whole geometry fold away (the kernels are partly issue-bound: 35-50 K warp instructions per tile);
// CI == 0: generic, geometry from the arguments.
template <int CI, int CO>
__global__ void __launch_bounds__(kAThreads, 1) nrf_a_kernel(const AArgs args_in) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  unsigned char* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  AArgs args = args_in;
  if (CI > 0) { args.ci = CI; args.co = CO; }
  const LevelGeom g = level_geom(args.ci, args.co);
  const LevelPack lp = level_pack(g);
  const int a_bytes = g.k1_steps * kStepBytes;                     // one x tile
  const int n_abuf = nrf_a_nbuf(g);
  const uint32_t kABStageBytes = (uint32_t)nrf_a_bstage_bytes(g);
  const uint32_t sm_a = smem_base, sm_b = smem_base + (uint32_t)(n_abuf * a_bytes);
  const uint32_t bars = sm_b + kABStages * kABStageBytes;
  auto bar = [&](int i) { return bars + (uint32_t)i * 8u; };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem_gen + (bars - smem_base) + kANumBars * 8);
  float* s_bias1 = reinterpret_cast<float*>(smem_gen + (bars - smem_base) + ((kANumBars * 8 + 16 + 15) & ~15));   // [n1p]: no global latency in the epilogue
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int HW = args.H * args.W;
  const int tiles_per_img = HW / kTile;
  const int n_tiles = args.n_img * tiles_per_img;
  for (int i = threadIdx.x; i < g.n1p; i += blockDim.x) s_bias1[i] = reinterpret_cast<const float*>(args.pack + lp.b1)[i];

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) { mbar_init(bar(kABarAFull + i), 8); mbar_init(bar(kABarAEmpty + i), 1); }
    for (int i = 0; i < kABStages; ++i) { mbar_init(bar(kABarBFull + i), 1); mbar_init(bar(kABarBEmpty + i), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(bar(kABarAccFull + i), 1); mbar_init(bar(kABarAccEmpty + i), 4); }
    fence_mbar_init();
  }
  if (warp == 13) tmem_alloc_512(smem_u32(const_cast<uint32_t*>(tmem_ptr_smem)));
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr_smem;
  // accumulator buffers: chunk j of a tile uses buffer (running chunk counter & 1) at TMEM column 0 / 256
  if (warp < 4) {
    // ======================================= epilogue: acc -> LReLU -> hi/lo -> t1 tiles in HBM; toRGB rows -> R_i ==========
    const int row = warp * 32 + lane;
    const float* bias1 = s_bias1;
    uint32_t cc = 0;   // running chunk counter
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int img = tile / tiles_per_img;
      const int p = (tile - img * tiles_per_img) * kTile + row;
      // Blur(up2(R_{i-1})) at this pixel, requested before the accumulators are awaited
      float ub[3] = {0.0f, 0.0f, 0.0f};
      if (args.rgb_prev != nullptr) {
        const int y = p / args.W, x = p - y * args.W;
#pragma unroll
        for (int j = 0; j < 3; ++j) ub[j] = up2_blur_at(args.rgb_prev + ((size_t)img * 3 + j) * (HW >> 2), args.H >> 1, args.W >> 1, y, x);
      }
      unsigned char* t1_tile = args.t1 + (size_t)tile * g.k2_steps * kStepBytes + (size_t)((row >> 3) * 256 + (row & 7) * 16);
      int n0 = 0;
      for (int ch = 0; ch < g.n_chunks; ++ch, ++cc) {
        const int buf = cc & 1;
        mbar_wait(bar(kABarAccFull + buf), (cc >> 1) & 1);
        tc_fence_after_sync();
        const uint32_t t_addr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(buf * 256);
        const int cn = g.chunk_n[ch];
        for (int c0 = 0; c0 < cn; c0 += 16) {
          uint32_t r[16];
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
              : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
              : "r"(t_addr + (uint32_t)c0)
              : "memory");
          float bv[16];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 b4 = reinterpret_cast<const float4*>(bias1 + n0 + c0)[q];
            bv[4 * q] = b4.x; bv[4 * q + 1] = b4.y; bv[4 * q + 2] = b4.z; bv[4 * q + 3] = b4.w;
          }
          tmem_wait_ld();
          if (c0 + 16 >= cn) {   // last group of this chunk read: the accumulator buffer can be refilled
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(kABarAccEmpty + buf));
          }
          const int col0 = n0 + c0;                 // global output column of r[0]; a multiple of 16 == K16 step of t1
          float v[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]) + bv[j];
          // toRGB rows (columns 2ci .. 2ci+2): R_i = toRGB_i(x) + b (+ Blur(up2(R_{i-1})))
          if (col0 + 16 > 2 * args.ci && col0 <= 2 * args.ci + 2) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int ch_rgb = col0 + j - 2 * args.ci;
              if (ch_rgb >= 0 && ch_rgb < 3) {
                args.rgb_out[((size_t)img * 3 + ch_rgb) * HW + p] = v[j] + (ch_rgb == 0 ? ub[0] : (ch_rgb == 1 ? ub[1] : ub[2]));
              }
            }
          }
          if (col0 < 2 * args.ci) {   // a K16 step of t1 (columns >= 2ci inside it are written as zeros)
            uint32_t hi[8], lo[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float a0 = (col0 + 2 * j < 2 * args.ci) ? lrelu02(v[2 * j]) : 0.0f;
              const float a1 = (col0 + 2 * j + 1 < 2 * args.ci) ? lrelu02(v[2 * j + 1]) : 0.0f;
              split2(a0, a1, hi[j], lo[j]);
            }
            unsigned char* dst = t1_tile + (size_t)(col0 >> 4) * kStepBytes;
            *reinterpret_cast<uint4*>(dst) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<uint4*>(dst + 128) = make_uint4(hi[4], hi[5], hi[6], hi[7]);
            *reinterpret_cast<uint4*>(dst + 4096) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            *reinterpret_cast<uint4*>(dst + 4096 + 128) = make_uint4(lo[4], lo[5], lo[6], lo[7]);
          }
        }
        n0 += cn;
      }
    }
  } else if (warp < 12) {
    // ======================================= loaders: x tile (optionally LReLU(Blur(src))) -> bf16 hi/lo A operand ==========
    // A sub-step = 8 of the 16 channels of a K16 step = one 16-byte chunk of the hi and of the lo operand of this thread's pixel row;
    // the two loader groups (4 warps each) share every tile, one K half each.  (Measured alternative, kept switchable: the groups
    // working on ALTERNATE tiles with both K halves per thread -- same bytes in flight per SM, 6 % slower.)
    const int row = ((warp - 4) & 3) * 32 + lane;
    const int grp = (warp - 4) >> 2;
    const bool split_tiles = false;
    const int n_sub = split_tiles ? 2 * g.k1_steps : g.k1_steps;
    auto sub_k = [&](int s_) { return split_tiles ? (s_ >> 1) : s_; };
    auto sub_kh = [&](int s_) { return split_tiles ? (s_ & 1) : grp; };
    // whole-row tiles (W a multiple of 128): a warp holds 32 consecutive x of one image row, so the horizontal taps of the separable
    // blur come from the neighbouring lanes (3 coalesced loads per channel instead of 9); lanes 0 / 31 fetch their outer column
    const bool fast = args.blur && (args.W % kTile == 0);
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      if (split_tiles && (int)(it & 1) != grp) continue;
      const int buf = (n_abuf == 2) ? (int)(it & 1) : 0;
      const uint32_t use = (n_abuf == 2) ? (it >> 1) : it;
      const int img = tile / tiles_per_img;
      const int p = (tile - img * tiles_per_img) * kTile + row;
      const int y = p / args.W, x = p - y * args.W;
      const int ym = reflect_idx(y - 1, args.H), yp = reflect_idx(y + 1, args.H);
      const int xm = reflect_idx(x - 1, args.W), xp = reflect_idx(x + 1, args.W);
      const int xe = lane == 0 ? xm : (lane == 31 ? xp : x);   // inner lanes re-read their own column (same cache line)
      const float* sp = args.src + (size_t)img * args.ci * HW;
      float* xo = args.x_out ? args.x_out + (size_t)img * args.ci * HW + p : nullptr;
      const size_t o0 = (size_t)ym * args.W, o1 = (size_t)y * args.W, o2 = (size_t)yp * args.W;
      mbar_wait(bar(kABarAEmpty + buf), (use & 1) ^ 1);
      const uint32_t a_row = sm_a + (uint32_t)(buf * a_bytes) + (uint32_t)((row >> 3) * 256 + (row & 7) * 16);
      // kLB K16 steps per memory round trip (their loads are all issued before the first value is consumed)
      constexpr int kLB = 2;
      auto emit = [&](int sub, const float (&vv)[8]) {   // one sub-step (8 channels) -> x_out, hi / lo chunks of the A operand
        const int k = sub_k(sub), kh = sub_kh(sub);
        if (xo != nullptr) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (k * 16 + kh * 8 + j < args.ci) xo[(size_t)(k * 16 + kh * 8 + j) * HW] = vv[j];
        }
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) split2(vv[2 * j], vv[2 * j + 1], hi[j], lo[j]);
        const uint32_t d = a_row + (uint32_t)(k * kStepBytes + kh * 128);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(d), "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]) : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(d + 4096), "r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3]) : "memory");
      };
      if (!args.blur) {
        // level 0 (one tile per CTA, nothing to overlap with): 6 K16 steps = 48 loads per thread per round trip
        constexpr int kL0 = 6;
        for (int k0 = 0; k0 < n_sub; k0 += kL0) {
          float v[kL0][8];
#pragma unroll
          for (int b = 0; b < kL0; ++b)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int c = sub_k(k0 + b) * 16 + sub_kh(k0 + b) * 8 + j;
              const float t = __ldg(sp + (size_t)min(c, args.ci - 1) * HW + p);   // unconditional load (clamped address): no branch, so
              v[b][j] = (c < args.ci) ? t : 0.0f;                                  // all loads of the batch are in flight together
            }
#pragma unroll
          for (int b = 0; b < kL0; ++b)
            if (k0 + b < n_sub) emit(k0 + b, v[b]);
        }
      }
      for (int k0 = 0; args.blur && k0 < n_sub; k0 += kLB) {
        float v[kLB][8];
        if (fast) {
          float m[kLB][8], e[kLB][8];
#pragma unroll
          for (int b = 0; b < kLB; ++b)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int c = sub_k(k0 + b) * 16 + sub_kh(k0 + b) * 8 + j;
              const float* pc = sp + (size_t)min(c, args.ci - 1) * HW;     // clamped: every load is unconditional (no branches)
              const bool ok = c < args.ci;
              const float t0 = __ldg(pc + o0 + x), t1v = __ldg(pc + o1 + x), t2v = __ldg(pc + o2 + x);
              const float e0 = __ldg(pc + o0 + xe), e1 = __ldg(pc + o1 + xe), e2 = __ldg(pc + o2 + xe);   // used by lanes 0 / 31 only
              m[b][j] = ok ? (t0 + 2.0f * t1v + t2v) : 0.0f;
              e[b][j] = ok ? (e0 + 2.0f * e1 + e2) : 0.0f;
            }
#pragma unroll
          for (int b = 0; b < kLB; ++b)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float l = __shfl_up_sync(0xffffffffu, m[b][j], 1), r = __shfl_down_sync(0xffffffffu, m[b][j], 1);
              if (lane == 0) l = e[b][j];
              if (lane == 31) r = e[b][j];
              v[b][j] = lrelu02((l + 2.0f * m[b][j] + r) * 0.0625f);
            }
        } else {
#pragma unroll
          for (int b = 0; b < kLB; ++b)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int c = sub_k(k0 + b) * 16 + sub_kh(k0 + b) * 8 + j;
              const float* pc = sp + (size_t)min(c, args.ci - 1) * HW;
              const float a = __ldg(pc + o0 + xm) + 2.0f * __ldg(pc + o0 + x) + __ldg(pc + o0 + xp);
              const float bb = __ldg(pc + o1 + xm) + 2.0f * __ldg(pc + o1 + x) + __ldg(pc + o1 + xp);
              const float d = __ldg(pc + o2 + xm) + 2.0f * __ldg(pc + o2 + x) + __ldg(pc + o2 + xp);
              v[b][j] = (c < args.ci) ? lrelu02((a + 2.0f * bb + d) * 0.0625f) : 0.0f;
            }
        }
#pragma unroll
        for (int b = 0; b < kLB; ++b)
          if (k0 + b < n_sub) emit(k0 + b, v[b]);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(kABarAFull + buf));
    }
  } else if (warp == 12) {
    // ======================================= weight TMA: W1 K16 slices of every chunk, once per tile ==========================
    if (lane == 0) {
      uint32_t slot = 0, phase = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const unsigned char* src = args.pack + lp.w1;
        for (int ch = 0; ch < g.n_chunks; ++ch) {
          const uint32_t bytes = (uint32_t)(g.chunk_n[ch] * 64);
          for (int k = 0; k < g.k1_steps; ++k) {
            mbar_wait_spin(bar(kABarBEmpty + slot), phase ^ 1);
            mbar_arrive_expect_tx(bar(kABarBFull + slot), bytes);
            bulk_g2s(sm_b + slot * kABStageBytes, src, bytes, bar(kABarBFull + slot));
            src += bytes;
            if (++slot == kABStages) { slot = 0; phase ^= 1; }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 13) {
    // ======================================= MMA issuer ==========================================================================
    // ONE elected lane runs the whole issue loop (waits included): a per-step elect + __syncwarp + whole-warp mbarrier wait costs more
    // than the 3 UMMAs of a K16 step take, and the tensor pipe only runs at rate when the UMMAs are issued back to back.
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    if (elect_one()) {
      uint32_t sb = 0, pb = 0, cc = 0, it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const int buf = (n_abuf == 2) ? (int)(it & 1) : 0;
        const uint32_t use = (n_abuf == 2) ? (it >> 1) : it;
        mbar_wait_spin(bar(kABarAFull + buf), use & 1);
        tc_fence_after_sync();
        const uint32_t a0 = (((sm_a + (uint32_t)(buf * a_bytes)) >> 4) & 0x3FFFu) | kDescLoLboNo;
        for (int ch = 0; ch < g.n_chunks; ++ch, ++cc) {
          const int ab = cc & 1;
          mbar_wait_spin(bar(kABarAccEmpty + ab), ((cc >> 1) & 1) ^ 1);
          tc_fence_after_sync();
          const uint32_t d = tmem_u + (uint32_t)(ab * 256);
          const uint32_t idesc = umma_idesc_bf16(128, g.chunk_n[ch]);
          const uint32_t slice_u = (uint32_t)((g.chunk_n[ch] * 32) >> 4);
          for (int k = 0; k < g.k1_steps; ++k) {
            mbar_wait_spin(bar(kABarBFull + sb), pb);
            tc_fence_after_sync();
            const uint32_t b0 = (((sm_b + sb * kABStageBytes) >> 4) & 0x3FFFu) | kDescLoLboNo;
            const uint64_t a_hi = mk_desc(a0 + (uint32_t)((k * kStepBytes) >> 4), kDescHiNoSw);
            const uint64_t a_lo = mk_desc(a0 + (uint32_t)((k * kStepBytes + 4096) >> 4), kDescHiNoSw);
            const uint64_t b_hi = mk_desc(b0, kDescHiNoSw), b_lo = mk_desc(b0 + slice_u, kDescHiNoSw);
            umma_ss(d, a_hi, b_hi, idesc, k == 0 ? 0u : 1u);
            umma_ss(d, a_lo, b_hi, idesc, 1u);
            umma_ss(d, a_hi, b_lo, idesc, 1u);
            umma_commit(bar(kABarBEmpty + sb));
            if (++sb == kABStages) { sb = 0; pb ^= 1; }
          }
          umma_commit(bar(kABarAccFull + ab));
          if (ch == g.n_chunks - 1) umma_commit(bar(kABarAEmpty + buf));
        }
      }
    }
    __syncwarp();
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 13) {
    tc_fence_after_sync();
    tmem_dealloc_512(tmem_base);
  }
}

static inline int nrf_a_smem_bytes(const LevelGeom& g) {
  return nrf_a_nbuf(g) * g.k1_steps * kStepBytes + kABStages * nrf_a_bstage_bytes(g) + kANumBars * 8 + 64 + 1024 + (g.n1p + 32) * 4;
}

// =====================================================================================================================
//  nrf_b_kernel
// =====================================================================================================================
struct BArgs {
  const unsigned char* t1;    // [tile][k2_steps][hi | lo]
  const unsigned char* pack;
  const float* xres;          // the level's input x [N][ci][HW] (PSU residual: + x.repeat(1,4,1,1))
  float* pre;                 // [N][co][2H][2W] = W3 pixel_shuffle(t2) + b3  (un-blurred, un-activated)
  int ci, co, H, W, n_img;
};

constexpr int kBASteps = 2;                                   // K16 steps of t1 per A-ring stage
constexpr int kBAStageBytes = kBASteps * kStepBytes;          // 16 KB
constexpr int kBMaxAStages = 8, kBMaxBStages = 8;
constexpr int kBA3Slots = 6;                                  // ring of K16 steps of the W3 GEMM's A operand (produced by the drain)
constexpr int kBThreads = 352;                                // warps 0-7 drain, 8 weight TMA, 9 MMA, 10 t1 TMA
constexpr int kBBarAFull = 0, kBBarAEmpty = kBMaxAStages, kBBarBFull = 2 * kBMaxAStages, kBBarBEmpty = kBBarBFull + kBMaxBStages,
              kBBarT2Full = kBBarBEmpty + kBMaxBStages, kBBarA3Ready = kBBarT2Full + 1, kBBarA3Free = kBBarA3Ready + kBA3Slots,
              kBBarPreFull = kBBarA3Free + kBA3Slots, kBBarPreFree = kBBarPreFull + 1, kBNumBars = kBBarPreFree + 1;
constexpr int kSmemBudget = 227 * 1024;

// weight ring stage: one K16 slice of the W2 group ([n2 rows x 32 B] hi + lo) or `w3_steps_per_stage` K16 slices of W3
__host__ __device__ inline int nrf_b_bstage_bytes(const LevelGeom& g) { return ((g.n2 > g.cop ? g.n2 : g.cop) * 64 + 127) & ~127; }
__host__ __device__ inline int nrf_b_w3_steps_per_stage(const LevelGeom& g) {
  int n = nrf_b_bstage_bytes(g) / (g.cop * 64);
  return n > g.k3_steps ? g.k3_steps : n;
}
__host__ __device__ inline int nrf_b_a3_slots(const LevelGeom& g) { return g.n2 / 16 < kBA3Slots ? g.n2 / 16 : kBA3Slots; }
__host__ __device__ inline int nrf_b_fixed_bytes(const LevelGeom& g) {
  return nrf_b_a3_slots(g) * kStepBytes + kBNumBars * 8 + 64 + 1024 + (4 * g.cip + g.cop + 64) * 4;   // W3 operand ring, barriers, biases
}
// ring depths from what is left: about two thirds for the weights (their slices are re-streamed for every 128-pixel tile and are
// 2x the bytes of the tile's t1), the rest for t1
__host__ __device__ inline int nrf_b_b_stages(const LevelGeom& g) {
  const int rem = kSmemBudget - nrf_b_fixed_bytes(g);
  int n = (rem * 17 / 25) / nrf_b_bstage_bytes(g);
  return n > kBMaxBStages ? kBMaxBStages : n;
}
__host__ __device__ inline int nrf_b_a_stages(const LevelGeom& g) {
  int n = (kSmemBudget - nrf_b_fixed_bytes(g) - nrf_b_b_stages(g) * nrf_b_bstage_bytes(g)) / kBAStageBytes;
  return n > kBMaxAStages ? kBMaxAStages : n;
}
static inline int nrf_b_smem_bytes(const LevelGeom& g) {
  return nrf_b_a_stages(g) * kBAStageBytes + nrf_b_b_stages(g) * nrf_b_bstage_bytes(g) + nrf_b_fixed_bytes(g);
}

// every level of this configuration fits the two kernels' smem / TMEM budgets
static inline bool level_supported(const LevelGeom& g) {
  if (g.n_chunks > kMaxChunks || g.chunk_n[0] > 256) return false;
  if (g.n2 > 288 || g.n2a > 256 || (g.n2b != 0 && g.n2b % 16 != 0)) return false;
  if ((g.n2 > 256 ? 320 : 256) + g.qg * g.cop > 512) return false;
  if (nrf_a_smem_bytes(g) > kSmemBudget || nrf_b_a_stages(g) < 2 || nrf_b_b_stages(g) < 2) return false;
  return true;
}

template <int CI, int CO>
__global__ void __launch_bounds__(kBThreads, 1) nrf_b_kernel(const BArgs args_in) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  unsigned char* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  BArgs args = args_in;
  if (CI > 0) { args.ci = CI; args.co = CO; }
  const LevelGeom g = level_geom(args.ci, args.co);
  const LevelPack lp = level_pack(g);
  const int kBAStages = nrf_b_a_stages(g), kBBStages = nrf_b_b_stages(g);
  const uint32_t kBBStageBytes = (uint32_t)nrf_b_bstage_bytes(g);
  const int w3_sps = nrf_b_w3_steps_per_stage(g);                 // K16 slices of W3 per ring stage
  const int w3_stages = (g.k3_steps + w3_sps - 1) / w3_sps;        // ring stages per pass over W3
  const bool w3_shared = (w3_stages == 1);                         // W3 fits ONE stage: loaded once per item, used by every sub-pixel
  const uint32_t sm_a = smem_base, sm_b = sm_a + (uint32_t)(kBAStages * kBAStageBytes), sm_a3 = sm_b + kBBStages * kBBStageBytes;
  const int a3_slots = nrf_b_a3_slots(g);
  const uint32_t bars = sm_a3 + (uint32_t)(a3_slots * kStepBytes);
  auto bar = [&](int i) { return bars + (uint32_t)i * 8u; };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem_gen + (bars - smem_base) + kBNumBars * 8);
  float* s_bias2 = reinterpret_cast<float*>(smem_gen + (bars - smem_base) + ((kBNumBars * 8 + 16 + 15) & ~15));   // [4 * cip] permuted (q, c)
  float* s_bias3 = s_bias2 + 4 * g.cip + 32;                                                        // [cop]
  for (int i = threadIdx.x; i < 4 * g.cip; i += blockDim.x) s_bias2[i] = reinterpret_cast<const float*>(args.pack + lp.b2)[i];
  for (int i = threadIdx.x; i < g.cop; i += blockDim.x) s_bias3[i] = reinterpret_cast<const float*>(args.pack + lp.b3)[i];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int HW = args.H * args.W;
  const int tiles_per_img = HW / kTile;
  const int n_groups = 4 / g.qg;
  const int n_items = args.n_img * tiles_per_img * n_groups;   // item = tile * n_groups + group (groups of a tile are adjacent: t1 stays in L2)
  const int a_stages_per_item = (g.k2_steps + kBASteps - 1) / kBASteps;
  const uint32_t pre_col = 256;   // TMEM: t2 accumulator at columns [0, n2) (n2 <= 272 -> second instruction at n2a), pre at [320, 320 + qg*cop)
  const uint32_t pre_base_col = (g.n2 > 256) ? 320u : pre_col;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kBMaxAStages; ++i) { mbar_init(bar(kBBarAFull + i), 1); mbar_init(bar(kBBarAEmpty + i), 1); }
    for (int i = 0; i < kBMaxBStages; ++i) { mbar_init(bar(kBBarBFull + i), 1); mbar_init(bar(kBBarBEmpty + i), 1); }
    mbar_init(bar(kBBarT2Full), 1);
    for (int i = 0; i < kBA3Slots; ++i) { mbar_init(bar(kBBarA3Ready + i), 4); mbar_init(bar(kBBarA3Free + i), 1); }
    mbar_init(bar(kBBarPreFull), 1);
    mbar_init(bar(kBBarPreFree), 8);
    fence_mbar_init();
  }
  if (warp == 9) tmem_alloc_512(smem_u32(const_cast<uint32_t*>(tmem_ptr_smem)));
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp < 8) {
    // ======================================= drain warps: thread = (row, column half) ==========================================
    const int half = warp >> 2;
    const int row = (warp & 3) * 32 + lane;
    const uint32_t t_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    const float* bias2 = s_bias2;
    const float* bias3 = s_bias3;
    const uint32_t a3_row = sm_a3 + (uint32_t)((row >> 3) * 256 + (row & 7) * 16);
    uint32_t it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int tile = item / n_groups, grp = item - tile * n_groups;
      const int img = tile / tiles_per_img;
      const int p = (tile - img * tiles_per_img) * kTile + row;
      const int h = p / args.W, w = p - h * args.W;
      const float* xr = args.xres + (size_t)img * args.ci * HW + p;
      // ---- t2 accumulator -> LReLU(acc + b2) + x[(4c + q) mod ci] -> hi/lo -> A operand of the W3 GEMM
      // K16 steps of the item's n2 columns are split between the two halves (even / odd steps).  The residual values do not depend
      // on the accumulator: those of step s + 2 are requested before step s is processed (and the first ones before the W2 GEMM
      // has finished), so their L2 latency never sits on the drain's critical path.
      const int n_steps = g.n2 / 16;
      auto load_res = [&](int s16, float (&rv)[16]) {
        const int col0 = s16 * 16;
        const int qq = col0 / g.cip, c0 = col0 - qq * g.cip;
        int cm = (4 * c0 + grp * g.qg + qq) % args.ci;       // source channel of the residual, advanced by 4 (mod ci) per column
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          rv[j] = __ldg(xr + (size_t)cm * HW);   // unconditional (cm is always a valid channel); columns >= ci are zeroed by the consumer
          cm += 4;
          if (cm >= args.ci) cm -= args.ci;
        }
      };
      auto drain_step = [&](int s16, const float (&rv)[16]) {
        const int col0 = s16 * 16;                 // column inside the item: sub-pixel qq = col0 / cip, channel c0 = col0 % cip
        const int qq = col0 / g.cip, c0 = col0 - qq * g.cip;
        const int q = grp * g.qg + qq;
        uint32_t r[16];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
              "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
            : "r"(t_lane + (uint32_t)col0)
            : "memory");
        float bv[16];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 b4 = reinterpret_cast<const float4*>(bias2 + q * g.cip + c0)[j];
          bv[4 * j] = b4.x; bv[4 * j + 1] = b4.y; bv[4 * j + 2] = b4.z; bv[4 * j + 3] = b4.w;
        }
        tmem_wait_ld();
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float a0 = (c0 + 2 * j < args.ci) ? lrelu02(__uint_as_float(r[2 * j]) + bv[2 * j]) + rv[2 * j] : 0.0f;
          const float a1 = (c0 + 2 * j + 1 < args.ci) ? lrelu02(__uint_as_float(r[2 * j + 1]) + bv[2 * j + 1]) + rv[2 * j + 1] : 0.0f;
          split2(a0, a1, hi[j], lo[j]);
        }
        // ring slot of this step: G = running step counter over the CTA's items (the W3 GEMM consumes the steps in order)
        const uint32_t G = it * (uint32_t)n_steps + (uint32_t)s16;
        const uint32_t slot = G % (uint32_t)a3_slots, use = G / (uint32_t)a3_slots;
        mbar_wait(bar(kBBarA3Free + slot), (use & 1) ^ 1);      // the MMAs that read the slot's previous contents have completed
        const uint32_t d = a3_row + slot * kStepBytes;
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(d), "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]) : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(d + 128), "r"(hi[4]), "r"(hi[5]), "r"(hi[6]), "r"(hi[7]) : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(d + 4096), "r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3]) : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(d + 4096 + 128), "r"(lo[4]), "r"(lo[5]), "r"(lo[6]), "r"(lo[7]) : "memory");
        fence_proxy_async_smem();      // generic-proxy st.shared -> visible to the tensor core
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(kBBarA3Ready + slot));   // 4 warps (this half's rows 0..127) complete the step
      };
      float rva[16], rvb[16];
      if (half < n_steps) load_res(half, rva);
      if (half + 2 < n_steps) load_res(half + 2, rvb);
      mbar_wait(bar(kBBarT2Full), it & 1);
      tc_fence_after_sync();
      for (int s16 = half; s16 < n_steps; s16 += 4) {
        drain_step(s16, rva);
        if (s16 + 4 < n_steps) load_res(s16 + 4, rva);
        if (s16 + 2 < n_steps) {
          drain_step(s16 + 2, rvb);
          if (s16 + 6 < n_steps) load_res(s16 + 6, rvb);
        }
      }
      // ---- pre accumulators -> + b3 -> pixel-shuffled store: pre[img][o][2h + qy][2w + qx]
      mbar_wait(bar(kBBarPreFull), it & 1);
      tc_fence_after_sync();
      {
        const size_t plane = (size_t)(4 * HW);
        const int W2 = 2 * args.W;
        float* pbase = args.pre + (size_t)img * args.co * plane + (size_t)(2 * h) * W2 + 2 * w;
        // columns of the pre accumulator: [qq][cop]; the two halves split the output channels (16-column groups)
        const int n_g16 = g.cop / 16;
        for (int g16 = half; g16 < n_g16; g16 += 2) {
          const int o0 = g16 * 16;
          float bv[16];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 b4 = reinterpret_cast<const float4*>(bias3 + o0)[j];
            bv[4 * j] = b4.x; bv[4 * j + 1] = b4.y; bv[4 * j + 2] = b4.z; bv[4 * j + 3] = b4.w;
          }
          if (g.qg == 4) {
            uint32_t r[4][16];
#pragma unroll
            for (int qq = 0; qq < 4; ++qq)
              asm volatile(
                  "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                  : "=r"(r[qq][0]), "=r"(r[qq][1]), "=r"(r[qq][2]), "=r"(r[qq][3]), "=r"(r[qq][4]), "=r"(r[qq][5]), "=r"(r[qq][6]),
                    "=r"(r[qq][7]), "=r"(r[qq][8]), "=r"(r[qq][9]), "=r"(r[qq][10]), "=r"(r[qq][11]), "=r"(r[qq][12]), "=r"(r[qq][13]),
                    "=r"(r[qq][14]), "=r"(r[qq][15])
                  : "r"(t_lane + pre_base_col + (uint32_t)(qq * g.cop + o0))
                  : "memory");
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              if (o0 + j < args.co) {
                float* o = pbase + (size_t)(o0 + j) * plane;
                *reinterpret_cast<float2*>(o) = make_float2(__uint_as_float(r[0][j]) + bv[j], __uint_as_float(r[1][j]) + bv[j]);
                *reinterpret_cast<float2*>(o + W2) = make_float2(__uint_as_float(r[2][j]) + bv[j], __uint_as_float(r[3][j]) + bv[j]);
              }
            }
          } else {
            uint32_t r[16];
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                  "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                : "r"(t_lane + pre_base_col + (uint32_t)o0)
                : "memory");
            tmem_wait_ld();
            const int q = grp;   // qg == 1: the item's only sub-pixel
            float* oq = pbase + (size_t)(q >> 1) * W2 + (q & 1);
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (o0 + j < args.co) oq[(size_t)(o0 + j) * plane] = __uint_as_float(r[j]) + bv[j];
          }
        }
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(kBBarPreFree));
    }
  } else if (warp == 10) {
    // ======================================= t1 TMA: the tile's K16 steps, kBASteps per stage ======================================
    if (lane == 0) {
      uint32_t slot = 0, phase = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int tile = item / n_groups;
        const unsigned char* src = args.t1 + (size_t)tile * g.k2_steps * kStepBytes;
        for (int s = 0; s < a_stages_per_item; ++s) {
          const int steps = min(kBASteps, g.k2_steps - s * kBASteps);
          const uint32_t bytes = (uint32_t)(steps * kStepBytes);
          mbar_wait_spin(bar(kBBarAEmpty + slot), phase ^ 1);
          mbar_arrive_expect_tx(bar(kBBarAFull + slot), bytes);
          bulk_g2s(sm_a + slot * kBAStageBytes, src, bytes, bar(kBBarAFull + slot));
          src += bytes;
          if (++slot == (uint32_t)kBAStages) { slot = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 8) {
    // ======================================= weight TMA: W2 slices of the item's group, then W3 slices once per sub-pixel =========
    if (lane == 0) {
      uint32_t slot = 0, phase = 0;
      const uint32_t bytes2 = (uint32_t)(g.n2 * 64), bytes3 = (uint32_t)(g.cop * 64);
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int grp = item % n_groups;
        const unsigned char* src = args.pack + lp.w2 + (size_t)grp * g.k2_steps * bytes2;
        for (int k = 0; k < g.k2_steps; ++k) {
          mbar_wait_spin(bar(kBBarBEmpty + slot), phase ^ 1);
          mbar_arrive_expect_tx(bar(kBBarBFull + slot), bytes2);
          bulk_g2s(sm_b + slot * kBBStageBytes, src, bytes2, bar(kBBarBFull + slot));
          src += bytes2;
          if (++slot == (uint32_t)kBBStages) { slot = 0; phase ^= 1; }
        }
        for (int qq = 0; qq < (w3_shared ? 1 : g.qg); ++qq) {
          src = args.pack + lp.w3;
          for (int s3 = 0; s3 < w3_stages; ++s3) {
            const int steps = min(w3_sps, g.k3_steps - s3 * w3_sps);
            const uint32_t bytes = (uint32_t)steps * bytes3;
            mbar_wait_spin(bar(kBBarBEmpty + slot), phase ^ 1);
            mbar_arrive_expect_tx(bar(kBBarBFull + slot), bytes);
            bulk_g2s(sm_b + slot * kBBStageBytes, src, bytes, bar(kBBarBFull + slot));
            src += bytes;
            if (++slot == (uint32_t)kBBStages) { slot = 0; phase ^= 1; }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 9) {
    // ======================================= MMA issuer (one elected lane runs the whole loop, see nrf_a_kernel) ====================
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    if (elect_one()) {
      uint32_t sa = 0, pa = 0, sb = 0, pb = 0, it = 0;
      const uint32_t idesc2a = umma_idesc_bf16(128, g.n2a), idesc2b = g.n2b ? umma_idesc_bf16(128, g.n2b) : 0u;
      const uint32_t idesc3 = umma_idesc_bf16(128, g.cop);
      const uint32_t slice2_u = (uint32_t)((g.n2 * 32) >> 4), slice3_u = (uint32_t)((g.cop * 32) >> 4);
      const uint32_t off2b_u = (uint32_t)((g.n2a * 32) >> 4);      // rows [n2a, n2) of a slice
      const uint32_t n_steps3 = (uint32_t)(g.n2 / 16);
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        // ---- W2 GEMM: t2[128 x n2] = t1[128 x 2ci] W2g^T   (the t2 accumulator is free: every K16 step of the previous item's drain
        //      was awaited by its W3 GEMM, which precedes this GEMM in program order)
        for (int k = 0; k < g.k2_steps; ++k) {
          const int kl = k % kBASteps;
          if (kl == 0) mbar_wait_spin(bar(kBBarAFull + sa), pa);
          mbar_wait_spin(bar(kBBarBFull + sb), pb);
          tc_fence_after_sync();
          const uint32_t a0 = (((sm_a + sa * kBAStageBytes + (uint32_t)(kl * kStepBytes)) >> 4) & 0x3FFFu) | kDescLoLboNo;
          const uint32_t b0 = (((sm_b + sb * kBBStageBytes) >> 4) & 0x3FFFu) | kDescLoLboNo;
          const uint64_t a_hi = mk_desc(a0, kDescHiNoSw), a_lo = mk_desc(a0 + (4096 >> 4), kDescHiNoSw);
          const uint32_t acc = k == 0 ? 0u : 1u;
          umma_ss(tmem_u, a_hi, mk_desc(b0, kDescHiNoSw), idesc2a, acc);
          umma_ss(tmem_u, a_lo, mk_desc(b0, kDescHiNoSw), idesc2a, 1u);
          umma_ss(tmem_u, a_hi, mk_desc(b0 + slice2_u, kDescHiNoSw), idesc2a, 1u);
          if (g.n2b) {
            const uint32_t d2 = tmem_u + (uint32_t)g.n2a;
            umma_ss(d2, a_hi, mk_desc(b0 + off2b_u, kDescHiNoSw), idesc2b, acc);
            umma_ss(d2, a_lo, mk_desc(b0 + off2b_u, kDescHiNoSw), idesc2b, 1u);
            umma_ss(d2, a_hi, mk_desc(b0 + slice2_u + off2b_u, kDescHiNoSw), idesc2b, 1u);
          }
          umma_commit(bar(kBBarBEmpty + sb));
          if (++sb == (uint32_t)kBBStages) { sb = 0; pb ^= 1; }
          if (kl == kBASteps - 1 || k == g.k2_steps - 1) {
            umma_commit(bar(kBBarAEmpty + sa));
            if (++sa == (uint32_t)kBAStages) { sa = 0; pa ^= 1; }
          }
        }
        umma_commit(bar(kBBarT2Full));
        // ---- W3 GEMM per sub-pixel: pre_q[128 x cop] = t2_q[128 x cip] W3^T, K16 step by K16 step as the drain produces them
        mbar_wait_spin(bar(kBBarPreFree), (it & 1) ^ 1);   // the previous item's pre accumulators have been read
        tc_fence_after_sync();
        for (int qq = 0; qq < g.qg; ++qq) {
          const uint32_t d = tmem_u + pre_base_col + (uint32_t)(qq * g.cop);
          for (int k = 0; k < g.k3_steps; ++k) {
            const uint32_t G = it * n_steps3 + (uint32_t)(qq * g.k3_steps + k);
            const uint32_t slot3 = G % (uint32_t)a3_slots, use3 = G / (uint32_t)a3_slots;
            const int kl = k % w3_sps;                       // slice inside the weight ring stage
            if (kl == 0 && !(w3_shared && qq > 0)) mbar_wait_spin(bar(kBBarBFull + sb), pb);
            mbar_wait_spin(bar(kBBarA3Ready + slot3), use3 & 1);
            tc_fence_after_sync();
            const bool last_of_stage = (kl == w3_sps - 1 || k == g.k3_steps - 1);
            const bool release = last_of_stage && (!w3_shared || qq == g.qg - 1);
            const uint32_t a0 = (((sm_a3 + slot3 * kStepBytes) >> 4) & 0x3FFFu) | kDescLoLboNo;
            const uint32_t b0 = (((sm_b + sb * kBBStageBytes + (uint32_t)(kl * g.cop * 64)) >> 4) & 0x3FFFu) | kDescLoLboNo;
            const uint64_t a_hi = mk_desc(a0, kDescHiNoSw), a_lo = mk_desc(a0 + (4096 >> 4), kDescHiNoSw);
            umma_ss(d, a_hi, mk_desc(b0, kDescHiNoSw), idesc3, k == 0 ? 0u : 1u);
            umma_ss(d, a_lo, mk_desc(b0, kDescHiNoSw), idesc3, 1u);
            umma_ss(d, a_hi, mk_desc(b0 + slice3_u, kDescHiNoSw), idesc3, 1u);
            umma_commit(bar(kBBarA3Free + slot3));
            if (release) {
              umma_commit(bar(kBBarBEmpty + sb));
              if (++sb == (uint32_t)kBBStages) { sb = 0; pb ^= 1; }
            }
          }
        }
        umma_commit(bar(kBBarPreFull));
      }
    }
    __syncwarp();
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after_sync();
    tmem_dealloc_512(tmem_base);
  }
}

}  // namespace nrf
}  // namespace gnrf
