// Weight-gradient GEMM on tcgen05 (sm_100a):   dW[n][k] = sum_img sum_p dY[img][n][p] * X[img][k][p]      (+ per-image bias grads)
//
// Backward of every 1x1 convolution on the path -- the radiance-MLP layers (models/mlp_nerf.py:95-119, nn.Conv2d(k=1) on
// [B,C,N_r,N_s]) and the neural-renderer convs (models/pixel_shuffle_upsample.py:26-38, models/neural_renderer.py:84-106) -- whose
// weight gradient is the contraction of the output gradient with the layer input over all pixels / sample points.
//
// UMMA view: M = 128 rows of dY (output channels), N = a chunk of <= 256 rows of X (input channels), K = pixels.  Both operands are
// stored channel-major with pixels contiguous, i.e. they are ALREADY K-major: a [rows x 64 px] fp32 block is converted to the bf16
// hi/lo pair and written as a SWIZZLE_128B K-major tile (row = channel, 128 B = 64 px).  Same bf16x3 split as the forward kernels
// (dy*x ~= dy_hi*x_hi + dy_lo*x_hi + dy_hi*x_lo, fp32 accumulate in TMEM).
//   * bias gradient for free: one extra all-ones "channel" is appended to X in shared memory (1.0 is exact in bf16), so column K of the
//     accumulator is sum_p dY[n][p].
//   * split-K over pixels: item = (image, 128-row tile of dY, X chunk, pixel range); every item writes its [128 x chunk] partial
//     accumulator to a workspace and wgrad_reduce_kernel sums them in a fixed order (deterministic, no atomics).
// Warp roles (288 threads, persistent, 1 CTA / SM): warps 0-7 convert (coalesced LDG.128 along pixels -> split -> swizzled
// st.shared) into a 2-stage ring; warp 8 issues the UMMAs; warps 0-3 drain the accumulator at the end of an item.
// HBM-bound by design (the operands are re-read once per tile/chunk pairing): algorithmic bytes are in DESIGN.md §3.5.
#include "common.cuh"
#include "tc_common.cuh"
#include "wgrad_tc.cuh"

namespace gnrf {
namespace tc {

constexpr int kWgStages = 2;
constexpr int kWgABytes = 32768;                 // [hi 16 KB | lo 16 KB], 128 rows
constexpr int kWgBBytes = 65536;                 // [hi 32 KB | lo 32 KB], <= 256 rows
constexpr int kWgStageBytes = kWgABytes + kWgBBytes;
constexpr int kWgThreads = 288;
constexpr int kWgConvWarps = 8;
constexpr int kWgMaxIter = (128 + 256) / 2 / kWgConvWarps;   // row pairs per converter warp and K-block
constexpr int kWgBarFull = 0, kWgBarEmpty = kWgStages, kWgBarAccFull = 2 * kWgStages, kWgBarAccEmpty = kWgBarAccFull + 1,
              kWgNumBars = kWgBarAccEmpty + 1;
constexpr int kWgSmemBars = kWgStages * kWgStageBytes;
constexpr int kWgSmemMisc = kWgSmemBars + kWgNumBars * 8;
constexpr int kWgSmemBytes = kWgSmemMisc + 64 + 1024;

struct WgradArgs {
  const float* dY; long long dy_img_stride;
  const float* X;  long long x_img_stride;
  int N_dy, K_x, rows_x;      // rows_x = K_x (+1 when the ones row is appended)
  int HW, n_img;
  int chunk_n, n_ch, n_mt, n_split, n_kb_total, n_items;
  float* partial;             // [n_items][128][chunk_n]
};

__global__ void __launch_bounds__(kWgThreads, 1) wgrad_tc_kernel(const WgradArgs args) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  unsigned char* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bars = smem_base + kWgSmemBars;
  auto bar = [&](int i) { return bars + (uint32_t)i * 8u; };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem_gen + kWgSmemMisc);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kWgStages; ++i) { mbar_init(bar(kWgBarFull + i), kWgConvWarps); mbar_init(bar(kWgBarEmpty + i), 1); }
    mbar_init(bar(kWgBarAccFull), 1);
    mbar_init(bar(kWgBarAccEmpty), 4);
    fence_mbar_init();
  }
  if (warp == 8) tmem_alloc_512(smem_u32(const_cast<uint32_t*>(tmem_ptr_smem)));
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr_smem;

  auto decode = [&](int item, int& img, int& mt, int& ch, int& kb0, int& kb1) {
    const int sp = item % args.n_split;
    int t = item / args.n_split;
    ch = t % args.n_ch; t /= args.n_ch;
    mt = t % args.n_mt;
    img = t / args.n_mt;
    kb0 = (int)(((long long)sp * args.n_kb_total) / args.n_split);
    kb1 = (int)(((long long)(sp + 1) * args.n_kb_total) / args.n_split);
  };

  if (warp < kWgConvWarps) {
    // ======================================= converters (+ epilogue on warps 0-3) =======================================
    const int half = lane >> 4, l16 = lane & 15;
    const int n_pairs = (128 + args.chunk_n) >> 1;
    const int n_iter = (n_pairs + kWgConvWarps - 1) / kWgConvWarps;
    uint32_t slot = 0, phase = 0;
    int it = 0;
    for (int item = blockIdx.x; item < args.n_items; item += gridDim.x, ++it) {
      int img, mt, ch, kb0, kb1;
      decode(item, img, mt, ch, kb0, kb1);
      const int m0 = mt * 128, n0 = ch * args.chunk_n;
      const float* dy_img = args.dY + (size_t)img * args.dy_img_stride;
      const float* x_img = args.X + (size_t)img * args.x_img_stride;
      for (int kb = kb0; kb < kb1; ++kb) {
        const int p = kb * 64 + l16 * 4;
        const bool p_ok = p < args.HW;
        // issue EVERY load of this K-block (<= 24 x 16 B per lane, ~96 KB per CTA in flight) before touching shared memory: the
        // loads do not depend on the ring, only the stores below wait for the stage to be released by the MMA warp
        float4 v[kWgMaxIter];
#pragma unroll
        for (int u = 0; u < kWgMaxIter; ++u) {
          const int row = (u * kWgConvWarps + warp) * 2 + half;
          v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (u < n_iter && p_ok) {
            if (row < 128) {
              if (m0 + row < args.N_dy) v[u] = __ldg(reinterpret_cast<const float4*>(dy_img + (size_t)(m0 + row) * args.HW + p));
            } else {
              const int n = n0 + row - 128;
              if (n < args.K_x) v[u] = __ldg(reinterpret_cast<const float4*>(x_img + (size_t)n * args.HW + p));
              else if (n < args.rows_x) v[u] = make_float4(1.f, 1.f, 1.f, 1.f);   // the ones row -> bias gradient column
            }
          }
        }
        mbar_wait(bar(kWgBarEmpty + slot), phase ^ 1);
        const uint32_t stage = smem_base + slot * kWgStageBytes;
#pragma unroll
        for (int u = 0; u < kWgMaxIter; ++u) {
          const int row = (u * kWgConvWarps + warp) * 2 + half;
          if (u < n_iter && row < 128 + args.chunk_n) {
            uint32_t h0, l0, h1, l1;
            split2(v[u].x, v[u].y, h0, l0);
            split2(v[u].z, v[u].w, h1, l1);
            const int r = row < 128 ? row : row - 128;
            const uint32_t base = stage + (row < 128 ? 0u : (uint32_t)kWgABytes) + a_row_offset(r) +
                                  ((((uint32_t)l16 >> 1) ^ ((uint32_t)r & 7u)) << 4) + ((uint32_t)l16 & 1u) * 8u;
            const uint32_t lo_off = row < 128 ? 16384u : 32768u;
            asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(base), "r"(h0), "r"(h1) : "memory");
            asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(base + lo_off), "r"(l0), "r"(l1) : "memory");
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(kWgBarFull + slot));
        if (++slot == kWgStages) { slot = 0; phase ^= 1; }
      }
      if (warp < 4) {
        // drain: thread = accumulator row (dY channel), 32 columns at a time -> partial[item][row][col]
        mbar_wait(bar(kWgBarAccFull), (uint32_t)(it & 1));
        tc_fence_after_sync();
        const int row = warp * 32 + lane;
        float* dst = args.partial + ((size_t)item * 128 + row) * args.chunk_n;
        const uint32_t t_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
        for (int c0 = 0; c0 < args.chunk_n; c0 += 32) {
          uint32_t r[32];
          tmem_ld32(t_addr + c0, r);
          tmem_wait_ld();
          const int ncols = min(32, args.chunk_n - c0);
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            if (j < ncols)
              *reinterpret_cast<float4*>(dst + c0 + j) = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                                                     __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
        }
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(kWgBarAccEmpty));
      }
    }
  } else {
    // ======================================= MMA issuer =======================================
    uint32_t slot = 0, phase = 0;
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t sbase_u = __shfl_sync(0xffffffffu, smem_base, 0);
    constexpr uint32_t kDescHiSw128 = (uint32_t)((1024 >> 4) | (1u << 14) | (2u << 29));
    constexpr uint32_t kDescLoLboSw = 1u << 16;
    auto mk = [](uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; };
    const uint32_t idesc = umma_idesc_bf16(128, args.chunk_n);
    int it = 0;
    for (int item = blockIdx.x; item < args.n_items; item += gridDim.x, ++it) {
      int img, mt, ch, kb0, kb1;
      decode(item, img, mt, ch, kb0, kb1);
      mbar_wait(bar(kWgBarAccEmpty), (uint32_t)((it & 1) ^ 1));
      tc_fence_after_sync();
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(bar(kWgBarFull + slot), phase);
        tc_fence_after_sync();
        const uint32_t a_hi0 = (((sbase_u + slot * kWgStageBytes) >> 4) & 0x3FFFu) | kDescLoLboSw;
        const uint32_t a_lo0 = a_hi0 + (16384 >> 4);
        const uint32_t b_hi0 = a_hi0 + (kWgABytes >> 4);
        const uint32_t b_lo0 = b_hi0 + (32768 >> 4);
        if (elect_one()) {
#pragma unroll
          for (int kl = 0; kl < 4; ++kl) {
            const uint64_t a_hi = mk(a_hi0 + (uint32_t)(kl * 2), kDescHiSw128);
            const uint64_t a_lo = mk(a_lo0 + (uint32_t)(kl * 2), kDescHiSw128);
            const uint64_t b_hi = mk(b_hi0 + (uint32_t)(kl * 2), kDescHiSw128);
            const uint64_t b_lo = mk(b_lo0 + (uint32_t)(kl * 2), kDescHiSw128);
            umma_ss(tmem_u, a_hi, b_hi, idesc, (kb == kb0 && kl == 0) ? 0u : 1u);
            umma_ss(tmem_u, a_lo, b_hi, idesc, 1u);
            umma_ss(tmem_u, a_hi, b_lo, idesc, 1u);
          }
          umma_commit(bar(kWgBarEmpty + slot));
        }
        __syncwarp();
        if (++slot == kWgStages) { slot = 0; phase ^= 1; }
      }
      if (elect_one()) umma_commit(bar(kWgBarAccFull));
      __syncwarp();
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after_sync();
    tmem_dealloc_512(tmem_base);
  }
}

// dW[n][k] (=|+=) sum over images and pixel splits of the partial accumulators; db_img[img][n] = column K_x summed over the splits.
__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, int N_dy, int K_x, int rows_x, int n_img, int n_mt, int n_ch,
                                    int n_split, int chunk_n, int accumulate, float* __restrict__ dW, float* __restrict__ db_img,
                                    int db_sum) {
  const long long total = (long long)n_mt * 128 * n_ch * chunk_n;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % chunk_n);
    long long t = i / chunk_n;
    const int ch = (int)(t % n_ch); t /= n_ch;
    const int m = (int)(t % 128);
    const int mt = (int)(t / 128);
    const int n = mt * 128 + m, k = ch * chunk_n + c;
    if (n >= N_dy || k >= rows_x) continue;
    float tot = 0.0f;
    for (int img = 0; img < n_img; ++img) {
      float s = 0.0f;
      const size_t item0 = (((size_t)img * n_mt + mt) * n_ch + ch) * n_split;
      for (int sp = 0; sp < n_split; ++sp) s += partial[((item0 + sp) * 128 + m) * chunk_n + c];
      if (k == K_x && db_img != nullptr && !db_sum) db_img[(size_t)img * N_dy + n] = s;
      tot += s;
    }
    if (k == K_x && db_img != nullptr && db_sum) db_img[n] = accumulate ? db_img[n] + tot : tot;
    if (k < K_x) {
      float* o = dW + (size_t)n * K_x + k;
      *o = accumulate ? *o + tot : tot;
    }
  }
}

WgradPlan wgrad_plan(int N_dy, int K_x, int n_img, int HW, bool want_bias) {
  WgradPlan pl;
  pl.rows_x = K_x + (want_bias ? 1 : 0);
  pl.n_mt = (N_dy + 127) / 128;
  pl.n_ch = (pl.rows_x + 255) / 256;
  int per = (pl.rows_x + pl.n_ch - 1) / pl.n_ch;
  pl.chunk_n = ((per + 15) / 16) * 16;
  pl.n_kb_total = (HW + 63) / 64;
  int groups = n_img * pl.n_mt * pl.n_ch;
  int sp = 148 / (groups > 0 ? groups : 1);
  if (sp < 1) sp = 1;
  if (sp > pl.n_kb_total) sp = pl.n_kb_total;
  pl.n_split = sp;
  pl.n_items = groups * sp;
  pl.partial_bytes = (size_t)pl.n_items * 128 * pl.chunk_n * sizeof(float);
  return pl;
}

int wgrad_tc_launch(const float* dY, long long dy_img_stride, const float* X, long long x_img_stride, int N_dy, int K_x, int n_img,
                    int HW, float* dW, float* db_img, int db_sum, int accumulate, void* ws, size_t ws_bytes, cudaStream_t st) {
  static int n_sm = 0;
  if (n_sm == 0) {
    int dev = 0;
    GNRF_CUDA(cudaGetDevice(&dev));
    GNRF_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    GNRF_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgSmemBytes));
  }
  if (HW % 4 != 0 || dy_img_stride % 4 != 0 || x_img_stride % 4 != 0)
    return fail(GNRF_ERR_ARG, "wgrad_tc: HW and image strides must be multiples of 4 floats (vector loads)");
  WgradPlan pl = wgrad_plan(N_dy, K_x, n_img, HW, db_img != nullptr);
  if (ws_bytes < pl.partial_bytes) return fail(GNRF_ERR_ARG, "wgrad_tc: workspace %zu < required %zu bytes", ws_bytes, pl.partial_bytes);
  WgradArgs a;
  a.dY = dY; a.dy_img_stride = dy_img_stride; a.X = X; a.x_img_stride = x_img_stride;
  a.N_dy = N_dy; a.K_x = K_x; a.rows_x = pl.rows_x; a.HW = HW; a.n_img = n_img;
  a.chunk_n = pl.chunk_n; a.n_ch = pl.n_ch; a.n_mt = pl.n_mt; a.n_split = pl.n_split; a.n_kb_total = pl.n_kb_total; a.n_items = pl.n_items;
  a.partial = static_cast<float*>(ws);
  int grid = pl.n_items < n_sm ? pl.n_items : n_sm;
  wgrad_tc_kernel<<<grid, kWgThreads, kWgSmemBytes, st>>>(a);
  long long total = (long long)pl.n_mt * 128 * pl.n_ch * pl.chunk_n;
  wgrad_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(a.partial, N_dy, K_x, pl.rows_x, n_img, pl.n_mt, pl.n_ch, pl.n_split,
                                                                     pl.chunk_n, accumulate, dW, db_img, db_sum);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(GNRF_ERR_CUDA, "wgrad_tc_launch: %s", cudaGetErrorString(e));
  count_launches(2);
  return GNRF_OK;
}

}  // namespace tc
}  // namespace gnrf

using namespace gnrf;

extern "C" size_t gnrf_wgrad_tc_workspace_bytes(int N, int K, int n_img, int HW) {
  if (N <= 0 || K <= 0 || n_img <= 0 || HW <= 0) return 0;
  const size_t a = tc::wgrad_plan(N, K, n_img, HW, true).partial_bytes, b = tc::wgrad_plan(N, K, n_img, HW, false).partial_bytes;
  return a > b ? a : b;
}

extern "C" int gnrf_wgrad_tc(const float* dY, long long dy_img_stride, const float* X, long long x_img_stride, int N, int K, int n_img,
                             int HW, float* dW, float* db, int db_sum, int accumulate, void* workspace, size_t workspace_bytes,
                             gnrf_stream_t stream) {
  GNRF_CHECK_ARG(dY && X && dW && workspace);
  GNRF_CHECK_ARG(N > 0 && K > 0 && n_img > 0 && HW > 0);
  return tc::wgrad_tc_launch(dY, dy_img_stride > 0 ? dy_img_stride : (long long)N * HW, X, x_img_stride > 0 ? x_img_stride : (long long)K * HW,
                             N, K, n_img, HW, dW, db, db_sum, accumulate, workspace, workspace_bytes, as_stream(stream));
}
