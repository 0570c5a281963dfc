// 1x1-convolution GEMM on tcgen05 for the 2-D neural renderer (sm_100a):   out[n][co][p] = epi( W[co][:] . X[n][:][p] + b[co] )
//
// Reference: the nn.Conv2d(k=1) layers of PixelShuffleUpsample (models/pixel_shuffle_upsample.py:26-38) and
// NeuralRenderer.feat_layers (models/neural_renderer.py:84-94, 103-104).
//
// UMMA view: M = 128 pixels (TMEM lane = pixel), N = a chunk of <= 256 output channels, K = input channels.
// Same bf16x3 split-precision scheme as the radiance MLP (x*w ~= x_hi*w_hi + x_lo*w_hi + x_hi*w_lo, fp32 accumulate).
// Warp-specialised, persistent (1 CTA / SM, 352 threads):
//   warps 0-3  epilogue  : tcgen05.ld accumulators -> +bias, LeakyReLU (+ PSU residual, pixel-shuffle scatter) -> fp32 NCHW stores
//                          (thread = pixel, so every store instruction writes 32 consecutive pixels of one channel: coalesced)
//   warp 10    X loader  : ONE TMA tensor load (cp.async.bulk.tensor.3d over a [n_img][K][HW] tensor map, zero fill outside) per
//                          [64 channels x 128 pixels] fp32 block of X into a 3-stage staging ring -- 96 KB in flight per SM, no
//                          registers, no per-thread load latency
//   warps 4-7  converters: read their pixel's 64 channel values from the staging block (conflict-free ld.shared), split to bf16
//                          hi/lo and write the UMMA-canonical SWIZZLE_128B K-major A tile into a 2-stage smem ring
//   warp  8    TMA       : streams the pre-arranged weight K-slices (W_hi then W_lo, no-swizzle core-matrix layout) with
//                          cp.async.bulk + mbarrier complete_tx through a 4-stage ring
//   warp  9    MMA       : issues 3 UMMAs per K16 step into one of TWO 256-column TMEM accumulator buffers, so the epilogue of
//                          work item i overlaps the MMAs of item i+1
#include <cuda.h>   // CUtensorMap (types only; the encoder is fetched through cudaGetDriverEntryPoint, no libcuda link)

#include "common.cuh"
#include "conv_tc.cuh"
#include "tc_common.cuh"

namespace gnrf {
namespace tc {

constexpr int kCvTile = 128;
constexpr int kCvAStages = 2, kCvAStageBytes = 32768;   // [hi 16 KB | lo 16 KB] of one K-block
constexpr int kCvBStages = 4, kCvBStageBytes = 16384;   // [W_hi slice | W_lo slice], chunk_n x 32 B each
constexpr int kCvSStages = 3, kCvSStageBytes = 32768;   // fp32 staging: [64 channels][128 px]
constexpr int kCvThreads = 352;
constexpr int kCvSmemA = 0;
constexpr int kCvSmemB = kCvAStages * kCvAStageBytes;
constexpr int kCvSmemS = kCvSmemB + kCvBStages * kCvBStageBytes;
constexpr int kCvSmemBars = kCvSmemS + kCvSStages * kCvSStageBytes;
constexpr int kCvBarAFull = 0, kCvBarAEmpty = kCvAStages, kCvBarBFull = 2 * kCvAStages, kCvBarBEmpty = kCvBarBFull + kCvBStages,
              kCvBarAccFull = kCvBarBEmpty + kCvBStages, kCvBarAccEmpty = kCvBarAccFull + 2, kCvBarSFull = kCvBarAccEmpty + 2,
              kCvBarSEmpty = kCvBarSFull + kCvSStages, kCvNumBars = kCvBarSEmpty + kCvSStages;
constexpr int kCvSmemMisc = kCvSmemBars + kCvNumBars * 8;
constexpr int kCvSmemBytes = kCvSmemMisc + 64 + 1024;

struct ConvArgs {
  const float* X;               // [n_img][K][HW]
  const unsigned char* wstream; // [chunk][k16][hi slice | lo slice]
  const float* bias;            // [n_chunks * chunk_n], zero padded
  float* out;
  const float* res;             // PSU residual source [n_img][Cres][HW]
  int Cres;
  int K, N, HW, Wd, n_img;
  int chunk_n, n_chunks, k16_steps, n_kb;
  int tiles_per_img, n_items, mode;
  int s_pitch, s_bytes;                     // staging row pitch (= TMA box width * 4) and box bytes
  // generic-GEMM extensions used by the training path (all optional; zero / null = the plain neural-renderer behaviour)
  long long x_img_stride, out_img_stride;   // elements between images of X / out (0 -> K*HW / N*HW)
  const float* bias_img;                    // [n_img][N] per-image bias added on top of the packed one (folded code columns)
  const float* mask; long long mask_img_stride; int mask_rows; float mask_slope;   // v *= mask>0 ? 1 : slope  (rows < mask_rows)
  const float* add;  long long add_img_stride;  int add_rows;                      // v += add                   (rows < add_rows)
};

__device__ __forceinline__ float lrelu02(float v) { return v >= 0.0f ? v : 0.2f * v; }

__global__ void __launch_bounds__(kCvThreads, 1) conv_tc_kernel(const ConvArgs args, const __grid_constant__ CUtensorMap x_map) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  unsigned char* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bars = smem_base + kCvSmemBars;
  auto bar = [&](int i) { return bars + (uint32_t)i * 8u; };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem_gen + kCvSmemMisc);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kCvAStages; ++i) { mbar_init(bar(kCvBarAFull + i), 4); mbar_init(bar(kCvBarAEmpty + i), 1); }
    for (int i = 0; i < kCvBStages; ++i) { mbar_init(bar(kCvBarBFull + i), 1); mbar_init(bar(kCvBarBEmpty + i), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(bar(kCvBarAccFull + i), 1); mbar_init(bar(kCvBarAccEmpty + i), 4); }
    for (int i = 0; i < kCvSStages; ++i) { mbar_init(bar(kCvBarSFull + i), 1); mbar_init(bar(kCvBarSEmpty + i), 4); }
    fence_mbar_init();
  }
  if (warp == 9) tmem_alloc_512(smem_u32(const_cast<uint32_t*>(tmem_ptr_smem)));
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const uint32_t stage_bytes_b = (uint32_t)(2 * args.chunk_n * 32);

  auto decode = [&](int item, int& img, int& tile, int& chunk) {
    chunk = item % args.n_chunks;
    int t = item / args.n_chunks;
    tile = t % args.tiles_per_img;
    img = t / args.tiles_per_img;
  };

  if (warp < 4) {
    // ======================================= epilogue =======================================
    const int row = warp * 32 + lane;
    int it = 0;
    for (int item = blockIdx.x; item < args.n_items; item += gridDim.x, ++it) {
      int img, tile, chunk;
      decode(item, img, tile, chunk);
      const int buf = it & 1;
      const int p = tile * kCvTile + row;
      const bool p_ok = p < args.HW;
      mbar_wait(bar(kCvBarAccFull + buf), (uint32_t)((it >> 1) & 1));
      tc_fence_after_sync();
      const uint32_t t_addr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(buf * 256);
      const int n0 = chunk * args.chunk_n;
      const int H = args.HW / args.Wd;
      const int h = p / args.Wd, w = p - h * args.Wd;
      const float* bias_c = args.bias + n0;               // zero padded to n_chunks * chunk_n
      const float* bias_i = args.bias_img ? args.bias_img + (size_t)img * args.N : nullptr;
      const size_t img_res = (size_t)img * args.Cres * args.HW + p;
      // act(v) = max(v, v * act_slope): 0.2 = LeakyReLU, 0 = ReLU, 1 = linear (branch-free)
      const float act_slope = (args.mode == CONV_EPI_LRELU || args.mode == CONV_EPI_PSU) ? 0.2f : (args.mode == CONV_EPI_RELU ? 0.0f : 1.0f);
      const bool extras = args.mask != nullptr || args.add != nullptr;
      for (int c0 = 0; c0 < args.chunk_n; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(t_addr + c0, r);
        // batch every global load of this 32-column group before any store, so their latencies overlap
        float bv[32], rv[32], av[32];
        const int ncols = min(32, args.chunk_n - c0);
        const int nvalid = min(ncols, args.N - n0 - c0);   // columns of this group that exist in the output (may be <= 0)
#pragma unroll
        for (int q = 0; q < 8; ++q) {   // the packed bias is zero padded 32 floats past the last chunk: unpredicated vector loads
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias_c + c0) + q);
          bv[4 * q] = b4.x; bv[4 * q + 1] = b4.y; bv[4 * q + 2] = b4.z; bv[4 * q + 3] = b4.w;
        }
        if (bias_i != nullptr) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < nvalid) bv[j] += __ldg(bias_i + n0 + c0 + j);
        }
        if (args.mode != CONV_EPI_PSU) {
          if (extras) {
            // raw mask / add values are only LOADED here (independent loads issue back to back); they are consumed after the wait
            const float* mk = args.mask ? args.mask + (size_t)img * args.mask_img_stride + (size_t)(n0 + c0) * args.HW + p : nullptr;
            const float* ad = args.add ? args.add + (size_t)img * args.add_img_stride + (size_t)(n0 + c0) * args.HW + p : nullptr;
            const int m_lim = (mk != nullptr && p_ok) ? min(nvalid, args.mask_rows - n0 - c0) : 0;
            const int a_lim = (ad != nullptr && p_ok) ? min(nvalid, args.add_rows - n0 - c0) : 0;
#pragma unroll
            for (int j = 0; j < 32; ++j) rv[j] = j < m_lim ? __ldg(mk + (size_t)j * args.HW) : 1.0f;
#pragma unroll
            for (int j = 0; j < 32; ++j) av[j] = j < a_lim ? __ldg(ad + (size_t)j * args.HW) : 0.0f;
          }
        } else {
          int cm = (n0 + c0) % args.Cres;   // (n % Cres), advanced incrementally
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            rv[j] = (p_ok && j < nvalid) ? __ldg(args.res + img_res + (size_t)cm * args.HW) : 0.0f;
            if (++cm == args.Cres) cm = 0;
          }
        }
        tmem_wait_ld();
        if (c0 + 32 >= args.chunk_n) {  // last group read: the accumulator buffer can be refilled
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar(kCvBarAccEmpty + buf));
        }
        if (!p_ok || nvalid <= 0) continue;
        if (args.mode != CONV_EPI_PSU) {
          float* o = args.out + (size_t)img * args.out_img_stride + (size_t)(n0 + c0) * args.HW + p;
          const size_t hw = (size_t)args.HW;
          if (!extras) {
            if (nvalid == 32) {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const float v = __uint_as_float(r[j]) + bv[j];
                *o = fmaxf(v, v * act_slope);
                o += hw;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                if (j < nvalid) {
                  const float v = __uint_as_float(r[j]) + bv[j];
                  *o = fmaxf(v, v * act_slope);
                }
                o += hw;
              }
            }
          } else {
            const float ms = args.mask_slope;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              if (j < nvalid) {
                float v = __uint_as_float(r[j]) + bv[j];
                v = fmaxf(v, v * act_slope);
                *o = fmaf(v, rv[j] > 0.0f ? 1.0f : ms, av[j]);
              }
              o += hw;
            }
          }
        } else {
          // + x.repeat(1,4,1,1), then pixel_shuffle(2): out[c][2h+i][2w+j] = in[4c+2i+j][h][w]  (pixel_shuffle_upsample.py:34-40)
          // n0 + c0 is a multiple of 16, so the 32 columns are 8 complete groups of 4 = (si, sj) in {0,1}^2 of channel c.
          const size_t plane = (size_t)(2 * H) * (2 * args.Wd);
          float* o = args.out + ((size_t)img * (args.N >> 2) + ((n0 + c0) >> 2)) * plane + (size_t)(2 * h) * (2 * args.Wd) + 2 * w;
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            if (4 * g < nvalid) {
              // the least significant mantissa bit of every stored value carries sign(pre-activation) (1 = negative): the backward
              // (psu_bwd_kernel) reads the LeakyReLU slope from it instead of recovering it from the cancellation-prone sh - x
              // (a 1-ulp, 6e-8 relative, perturbation of the forward value)
              auto psu = [](float z, float res) {
                const float o = lrelu02(z) + res;
                return __uint_as_float((__float_as_uint(o) & ~1u) | (z < 0.0f ? 1u : 0u));
              };
              float v0 = psu(__uint_as_float(r[4 * g + 0]) + bv[4 * g + 0], rv[4 * g + 0]);
              float v1 = psu(__uint_as_float(r[4 * g + 1]) + bv[4 * g + 1], rv[4 * g + 1]);
              float v2 = psu(__uint_as_float(r[4 * g + 2]) + bv[4 * g + 2], rv[4 * g + 2]);
              float v3 = psu(__uint_as_float(r[4 * g + 3]) + bv[4 * g + 3], rv[4 * g + 3]);
              *reinterpret_cast<float2*>(o) = make_float2(v0, v1);                       // row 2h,   cols 2w, 2w+1
              *reinterpret_cast<float2*>(o + 2 * args.Wd) = make_float2(v2, v3);         // row 2h+1
            }
            o += plane;
          }
        }
      }
    }
  } else if (warp < 8) {
    // ======================================= converters: staged fp32 [channel][pixel] -> bf16 hi/lo SW128 A tiles =================
    const int row = (warp - 4) * 32 + lane;
    uint32_t slot = 0, phase = 0, ss = 0, sphase = 0;
    for (int item = blockIdx.x; item < args.n_items; item += gridDim.x) {
      int img, tile, chunk;
      decode(item, img, tile, chunk);
      const int p = tile * kCvTile + row;
      const bool p_ok = p < args.HW;
      for (int kb = 0; kb < args.n_kb; ++kb) {
        mbar_wait(bar(kCvBarSFull + ss), sphase);
        const uint32_t s_addr = smem_base + kCvSmemS + ss * kCvSStageBytes + (uint32_t)row * 4u;
        const int nch = min(64, args.K - kb * 64);
        float v[64];
#pragma unroll
        for (int j = 0; j < 64; ++j) {
          float t;
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(t) : "r"(s_addr + (uint32_t)(j * args.s_pitch)));
          v[j] = (p_ok && j < nch) ? t : 0.0f;   // rows beyond the TMA box hold stale data (inside the box TMA zero-fills)
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(kCvBarSEmpty + ss));
        if (++ss == kCvSStages) { ss = 0; sphase ^= 1; }
        uint32_t hi[32], lo[32];
        split_row64(v, hi, lo);
        mbar_wait(bar(kCvBarAEmpty + slot), phase ^ 1);
        const uint32_t a_addr = smem_base + kCvSmemA + slot * kCvAStageBytes + a_row_offset(row);
        st_shared_row128(a_addr, (uint32_t)(row & 7), hi);
        st_shared_row128(a_addr + 16384, (uint32_t)(row & 7), lo);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(kCvBarAFull + slot));
        if (++slot == kCvAStages) { slot = 0; phase ^= 1; }
      }
    }
  } else if (warp == 10) {
    // ======================================= X loader: one TMA tensor load per K-block ===========================================
    if (lane == 0) {
      uint32_t ss = 0, sphase = 0;
      const uint64_t map_ptr = reinterpret_cast<uint64_t>(&x_map);
      for (int item = blockIdx.x; item < args.n_items; item += gridDim.x) {
        int img, tile, chunk;
        decode(item, img, tile, chunk);
        const int p0 = tile * kCvTile;
        for (int kb = 0; kb < args.n_kb; ++kb) {
          mbar_wait(bar(kCvBarSEmpty + ss), sphase ^ 1);
          mbar_arrive_expect_tx(bar(kCvBarSFull + ss), (uint32_t)args.s_bytes);
          asm volatile(
              "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                  smem_base + kCvSmemS + ss * kCvSStageBytes),
              "l"(map_ptr), "r"(p0), "r"(kb * 64), "r"(img), "r"(bar(kCvBarSFull + ss))
              : "memory");
          if (++ss == kCvSStages) { ss = 0; sphase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 8) {
    // ======================================= TMA producer (weights) =======================================
    if (lane == 0) {
      uint32_t slot = 0, phase = 0;
      for (int item = blockIdx.x; item < args.n_items; item += gridDim.x) {
        int img, tile, chunk;
        decode(item, img, tile, chunk);
        const unsigned char* src = args.wstream + (size_t)chunk * args.k16_steps * stage_bytes_b;
        for (int k = 0; k < args.k16_steps; ++k) {
          mbar_wait(bar(kCvBarBEmpty + slot), phase ^ 1);
          mbar_arrive_expect_tx(bar(kCvBarBFull + slot), stage_bytes_b);
          bulk_g2s(smem_base + kCvSmemB + slot * kCvBStageBytes, src, stage_bytes_b, bar(kCvBarBFull + slot));
          src += stage_bytes_b;
          if (++slot == kCvBStages) { slot = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 9) {
    // ======================================= MMA issuer (converged warp, one elected lane issues) =======================================
    uint32_t sa = 0, pa = 0, sb = 0, pb = 0;
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t sbase_u = __shfl_sync(0xffffffffu, smem_base, 0);
    constexpr uint32_t kDescHiSw128 = (uint32_t)((1024 >> 4) | (1u << 14) | (2u << 29));
    constexpr uint32_t kDescHiNoSw = (uint32_t)((256 >> 4) | (1u << 14));
    constexpr uint32_t kDescLoLboSw = 1u << 16, kDescLoLboNo = (128u >> 4) << 16;
    auto mk = [](uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; };
    const uint32_t idesc = umma_idesc_bf16(128, args.chunk_n);
    const uint32_t slice_u = (uint32_t)((args.chunk_n * 32) >> 4);
    int it = 0;
    for (int item = blockIdx.x; item < args.n_items; item += gridDim.x, ++it) {
      const int buf = it & 1;
      mbar_wait(bar(kCvBarAccEmpty + buf), (uint32_t)(((it >> 1) & 1) ^ 1));
      tc_fence_after_sync();
      const uint32_t d = tmem_u + (uint32_t)(buf * 256);
      for (int kb = 0; kb < args.n_kb; ++kb) {
        mbar_wait(bar(kCvBarAFull + sa), pa);
        tc_fence_after_sync();
        const uint32_t a_hi0 = (((sbase_u + kCvSmemA + sa * kCvAStageBytes) >> 4) & 0x3FFFu) | kDescLoLboSw;
        const uint32_t a_lo0 = a_hi0 + (16384 >> 4);
        for (int kl = 0; kl < 4; ++kl) {
          if (kb * 4 + kl >= args.k16_steps) break;
          mbar_wait(bar(kCvBarBFull + sb), pb);
          tc_fence_after_sync();
          const uint32_t b_hi0 = (((sbase_u + kCvSmemB + sb * kCvBStageBytes) >> 4) & 0x3FFFu) | kDescLoLboNo;
          if (elect_one()) {
            const uint64_t a_hi = mk(a_hi0 + (uint32_t)(kl * 2), kDescHiSw128);
            const uint64_t a_lo = mk(a_lo0 + (uint32_t)(kl * 2), kDescHiSw128);
            const uint64_t b_hi = mk(b_hi0, kDescHiNoSw);
            const uint64_t b_lo = mk(b_hi0 + slice_u, kDescHiNoSw);
            umma_ss(d, a_hi, b_hi, idesc, (kb == 0 && kl == 0) ? 0u : 1u);
            umma_ss(d, a_lo, b_hi, idesc, 1u);
            umma_ss(d, a_hi, b_lo, idesc, 1u);
            umma_commit(bar(kCvBarBEmpty + sb));
          }
          __syncwarp();
          if (++sb == kCvBStages) { sb = 0; pb ^= 1; }
        }
        if (elect_one()) umma_commit(bar(kCvBarAEmpty + sa));
        __syncwarp();
        if (++sa == kCvAStages) { sa = 0; pa ^= 1; }
      }
      if (elect_one()) umma_commit(bar(kCvBarAccFull + buf));
      __syncwarp();
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after_sync();
    tmem_dealloc_512(tmem_base);
  }
}

// Weight stream of one layer: [chunk][k16][half][ (row/8)*16 + k_half*8 + (row%8) ] 16-byte chunks; bias zero-padded.
// W(n,k) = W[n * sn + k * sk]  (sn = K, sk = 1: dense [N][K]; sn = 1, sk = N: the transpose of a dense [K][N] matrix).
__global__ void conv_pack_kernel(const float* __restrict__ W, long long sn, long long sk, const float* __restrict__ b, int N, int K,
                                 int chunk_n, int n_chunks, int k16_steps, unsigned char* __restrict__ stream,
                                 float* __restrict__ bias_out) {
  const size_t chunks_per_slice = (size_t)chunk_n * 2;
  const size_t total = (size_t)n_chunks * k16_steps * 2 * chunks_per_slice;
  for (size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x; c < total; c += (size_t)gridDim.x * blockDim.x) {
    size_t rem = c % chunks_per_slice;
    size_t s = c / chunks_per_slice;
    const int half = (int)(s & 1); s >>= 1;
    const int k16 = (int)(s % k16_steps);
    const int chunk = (int)(s / k16_steps);
    const int row = (int)(rem >> 4) * 8 + (int)(rem & 7);
    const int k_half = (int)(rem >> 3) & 1;
    const int n = chunk * chunk_n + row;
    const int k0 = k16 * 16 + k_half * 8;
    uint32_t out[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int ka = k0 + 2 * q, kb = ka + 1;
      float a = (n < N && ka < K) ? W[(size_t)n * sn + (size_t)ka * sk] : 0.0f;
      float bb = (n < N && kb < K) ? W[(size_t)n * sn + (size_t)kb * sk] : 0.0f;
      uint32_t hi, lo;
      split2(a, bb, hi, lo);
      out[q] = half ? lo : hi;
    }
    *reinterpret_cast<uint4*>(stream + c * 16) = make_uint4(out[0], out[1], out[2], out[3]);
  }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_chunks * chunk_n + 32; i += gridDim.x * blockDim.x)
    bias_out[i] = (i < N && b != nullptr) ? b[i] : 0.0f;
}

ConvLayerPlan conv_layer_plan(int N, int K) {
  ConvLayerPlan pl;
  pl.N = N;
  pl.K = K;
  pl.n_chunks = (N + 255) / 256;
  int per = (N + pl.n_chunks - 1) / pl.n_chunks;
  pl.chunk_n = ((per + 15) / 16) * 16;
  pl.k16_steps = (K + 15) / 16;
  pl.n_kb = (K + 63) / 64;
  pl.stream_bytes = (size_t)pl.n_chunks * pl.k16_steps * 2 * pl.chunk_n * 32;
  pl.bias_floats = (size_t)pl.n_chunks * pl.chunk_n + 32;   // + 32: the epilogue reads whole 32-float groups
  pl.total_bytes = ((pl.stream_bytes + pl.bias_floats * sizeof(float)) + 255) & ~(size_t)255;
  return pl;
}

int conv_tc_pack(const ConvLayerPlan& pl, const float* W, const float* b, unsigned char* dst, cudaStream_t st) {
  return conv_tc_pack_strided(pl, W, pl.K, 1, b, dst, st);
}

int conv_tc_pack_strided(const ConvLayerPlan& pl, const float* W, long long sn, long long sk, const float* b, unsigned char* dst,
                         cudaStream_t st) {
  float* bias_out = reinterpret_cast<float*>(dst + pl.stream_bytes);
  conv_pack_kernel<<<148, 256, 0, st>>>(W, sn, sk, b, pl.N, pl.K, pl.chunk_n, pl.n_chunks, pl.k16_steps, dst, bias_out);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(GNRF_ERR_CUDA, "conv_tc_pack: %s", cudaGetErrorString(e));
  count_launches(1);
  return GNRF_OK;
}

typedef CUresult (*TensorMapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                      const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static TensorMapEncodeFn tensor_map_encoder() {
  static TensorMapEncodeFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<TensorMapEncodeFn>(p);
  }
  return fn;
}

int conv_tc_launch(const ConvLayerPlan& pl, const unsigned char* packed, const float* X, float* out, const float* res, int Cres,
                   int n_img, int HW, int Wd, int mode, cudaStream_t st) {
  ConvExtras ex = {};
  return conv_tc_launch_ex(pl, packed, X, out, res, Cres, n_img, HW, Wd, mode, ex, st);
}

int conv_tc_launch_ex(const ConvLayerPlan& pl, const unsigned char* packed, const float* X, float* out, const float* res, int Cres,
                      int n_img, int HW, int Wd, int mode, const ConvExtras& ex, cudaStream_t st) {
  int n_sm = 0;
  {
    int rc = device_once(kOnceConvTc, &n_sm, []() -> int {
      GNRF_CUDA(cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kCvSmemBytes));
      return GNRF_OK;
    });
    if (rc != GNRF_OK) return rc;
  }
  ConvArgs a;
  a.X = X;
  a.wstream = packed;
  a.bias = reinterpret_cast<const float*>(packed + pl.stream_bytes);
  a.out = out;
  a.res = res;
  a.Cres = Cres > 0 ? Cres : 1;
  a.K = pl.K; a.N = pl.N; a.HW = HW; a.Wd = Wd; a.n_img = n_img;
  a.chunk_n = pl.chunk_n; a.n_chunks = pl.n_chunks; a.k16_steps = pl.k16_steps; a.n_kb = pl.n_kb;
  a.tiles_per_img = (HW + kCvTile - 1) / kCvTile;
  a.n_items = n_img * a.tiles_per_img * pl.n_chunks;
  a.mode = mode;
  if (HW % 4 != 0 || (ex.x_img_stride % 4) != 0 || (reinterpret_cast<uintptr_t>(X) & 15) != 0)
    return fail(GNRF_ERR_ARG, "conv_tc_launch: X must be 16-byte aligned with HW and the image stride multiples of 4 floats (TMA rows)");
  a.x_img_stride = ex.x_img_stride > 0 ? ex.x_img_stride : (long long)pl.K * HW;
  a.out_img_stride = ex.out_img_stride > 0 ? ex.out_img_stride : (long long)pl.N * HW;
  a.bias_img = ex.bias_img;
  a.mask = ex.mask; a.mask_img_stride = ex.mask_img_stride > 0 ? ex.mask_img_stride : (long long)pl.N * HW;
  a.mask_rows = ex.mask_rows > 0 ? ex.mask_rows : pl.N; a.mask_slope = ex.mask_slope;
  a.add = ex.add; a.add_img_stride = ex.add_img_stride > 0 ? ex.add_img_stride : (long long)pl.N * HW;
  a.add_rows = ex.add_rows > 0 ? ex.add_rows : pl.N;
  // fp32 tensor [n_img][K][HW] (strides in bytes), box = [1][<=64 channels][<=128 px]; out-of-range elements are zero-filled
  TensorMapEncodeFn enc = tensor_map_encoder();
  if (enc == nullptr) return fail(GNRF_ERR_CUDA, "conv_tc_launch: cuTensorMapEncodeTiled is not available from the driver");
  CUtensorMap x_map;
  const cuuint64_t gdim[3] = {(cuuint64_t)HW, (cuuint64_t)pl.K, (cuuint64_t)n_img};
  const cuuint64_t gstr[2] = {(cuuint64_t)HW * 4, (cuuint64_t)a.x_img_stride * 4};
  const cuuint32_t box[3] = {(cuuint32_t)(HW < kCvTile ? HW : kCvTile), (cuuint32_t)(pl.K < 64 ? pl.K : 64), 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  CUresult cr = enc(&x_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(X), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) return fail(GNRF_ERR_CUDA, "conv_tc_launch: cuTensorMapEncodeTiled failed (%d) for HW=%d K=%d n_img=%d", (int)cr, HW, pl.K, n_img);
  a.s_pitch = (int)box[0] * 4;
  a.s_bytes = (int)(box[0] * box[1]) * 4;
  int grid = a.n_items < n_sm ? a.n_items : n_sm;
  conv_tc_kernel<<<grid, kCvThreads, kCvSmemBytes, st>>>(a, x_map);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(GNRF_ERR_CUDA, "conv_tc_launch: %s", cudaGetErrorString(e));
  count_launches(1);
  return GNRF_OK;
}

}  // namespace tc
}  // namespace gnrf

using namespace gnrf;

// ---- generic entry points (training path): any 1x1 conv / per-point Linear forward, and its input gradient with W^T packed ----
extern "C" size_t gnrf_conv_tc_packed_bytes(int N, int K) {
  if (N <= 0 || K <= 0) return 0;
  return tc::conv_layer_plan(N, K).total_bytes;
}

extern "C" int gnrf_conv_tc_pack(const float* W, const float* bias, int N, int K, int transposed, void* packed, gnrf_stream_t stream) {
  GNRF_CHECK_ARG(W && packed && N > 0 && K > 0);
  GNRF_CHECK_ARG((reinterpret_cast<uintptr_t>(packed) & 127) == 0);
  // transposed: W points at a dense [K][N] matrix (a forward weight [out = K][in = N]) and the GEMM uses its transpose
  return tc::conv_tc_pack_strided(tc::conv_layer_plan(N, K), W, transposed ? 1 : K, transposed ? N : 1, bias,
                                  static_cast<unsigned char*>(packed), as_stream(stream));
}

extern "C" int gnrf_conv_tc(const void* packed, int N, int K, const float* X, long long x_img_stride, const float* bias_img, float* out,
                            long long out_img_stride, int act, const float* mask, long long mask_img_stride, int mask_rows,
                            float mask_slope, const float* add, long long add_img_stride, int add_rows, int n_img, int HW,
                            gnrf_stream_t stream) {
  GNRF_CHECK_ARG(packed && X && out && N > 0 && K > 0 && n_img > 0 && HW > 0);
  GNRF_CHECK_ARG(act == GNRF_ACT_NONE || act == GNRF_ACT_RELU || act == GNRF_ACT_LRELU02);
  tc::ConvExtras ex = {};
  ex.x_img_stride = x_img_stride; ex.out_img_stride = out_img_stride; ex.bias_img = bias_img;
  ex.mask = mask; ex.mask_img_stride = mask_img_stride; ex.mask_rows = mask_rows; ex.mask_slope = mask_slope;
  ex.add = add; ex.add_img_stride = add_img_stride; ex.add_rows = add_rows;
  const int mode = act == GNRF_ACT_NONE ? tc::CONV_EPI_LINEAR : (act == GNRF_ACT_RELU ? tc::CONV_EPI_RELU : tc::CONV_EPI_LRELU);
  return tc::conv_tc_launch_ex(tc::conv_layer_plan(N, K), static_cast<const unsigned char*>(packed), X, out, nullptr, 1, n_img, HW, HW,
                               mode, ex, as_stream(stream));
}
