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
#include <cuda.h>   // CUtensorMap (types only; the encoder comes through cudaGetDriverEntryPoint)

#include "common.cuh"
#include "tc_common.cuh"
#include "wgrad_tc.cuh"

namespace gnrf {
namespace tc {

constexpr int kWgTileBytes = 98304;              // bf16 operand tiles of one K-block: A [hi 16 KB | lo 16 KB] + B [hi 32 KB | lo 32 KB]
constexpr int kWgABytes = 32768;
constexpr int kWgStagBytes = 49152;              // fp32 staging of HALF a K-block: (128 + 256) rows x 32 px x 4 B; two slots
constexpr int kWgThreads = 320;                  // warps 0-7 converters (0-3 also drain), 8 MMA, 9 TMA loader
constexpr int kWgConvWarps = 8;
constexpr int kWgMaxIter = (128 + 256) / 4 / kWgConvWarps;   // 4-row groups per converter warp and half K-block (12)
constexpr int kWgBarStgFull = 0, kWgBarStgEmpty = 2, kWgBarTileFull = 4, kWgBarTileEmpty = 5, kWgBarAccFull = 6, kWgBarAccEmpty = 7,
              kWgNumBars = 8;
constexpr int kWgSmemStag = kWgTileBytes;
constexpr int kWgSmemBars = kWgTileBytes + 2 * kWgStagBytes;
constexpr int kWgSmemMisc = kWgSmemBars + kWgNumBars * 8;
constexpr int kWgSmemBytes = kWgSmemMisc + 64 + 1024;

struct WgradArgs {
  int N_dy, K_x, rows_x;      // rows_x = K_x (+1 when the ones row is appended)
  int HW, n_img;
  int chunk_n, n_ch, n_mt, n_split, n_kb_total, n_items;
  int pitch;                  // staging row pitch in bytes (= TMA box width * 4)
  int rows_dy_box, rows_x_box, stage_tx_bytes;
  float* partial;             // [n_items][128][chunk_n]
};

// Data path: a K-block (64 pixels) arrives as two HALF K-blocks (32 pixels).  For each half the loader warp issues two TMA tensor
// loads (dY rows of the 128-row tile, X rows of the chunk; zero fill outside the tensors) into one of two fp32 staging slots; the
// converter warps read a slot (ld.shared.v4), split to bf16 hi/lo in registers and release the slot at once -- so the TMA of the next
// half is always in flight while the current one is converted (with a single staging block, or with per-thread LDGs, the load round
// trip and the conversion were serialised: 5.9 K cycles per K-block against 1.25 K of UMMA work; profiles/r1_wgrad_tc_kernel.md).
// When both halves are in registers the converters wait for the MMA warp to release the operand tiles and write them.
__global__ void __launch_bounds__(kWgThreads, 1) wgrad_tc_kernel(const WgradArgs args, const __grid_constant__ CUtensorMap dy_map,
                                                                 const __grid_constant__ CUtensorMap x_map) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  unsigned char* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bars = smem_base + kWgSmemBars;
  auto bar = [&](int i) { return bars + (uint32_t)i * 8u; };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem_gen + kWgSmemMisc);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) { mbar_init(bar(kWgBarStgFull + i), 1); mbar_init(bar(kWgBarStgEmpty + i), kWgConvWarps); }
    mbar_init(bar(kWgBarTileFull), kWgConvWarps);
    mbar_init(bar(kWgBarTileEmpty), 1);
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
    // Row assignment: in iteration u, the 8 lanes of quarter q = lane / 8 of warp w handle tile row  u*32 + lane_row  (lane_row < 32,
    // see below), 4 pixels (16 B of the 128-B staging row) per lane: rows 0..127 (u < 4) are the dY tile, rows 128.. (u >= 4) the X
    // chunk.  Everything that does not change from K-block to K-block is hoisted: per-iteration validity as bit masks (per item),
    // shared-memory addresses as lane constants + u * (compile-time step).
    // The four rows of one warp instruction are r0, r0+2, r0+4, r0+6: with the SWIZZLE_128B chunk XOR (row & 7) the 64-byte half rows
    // written by st.shared then fall into disjoint bank halves pairwise (2 wavefronts per 256 B; consecutive rows collided 4-way:
    // ncu showed 50 % of all shared-memory wavefronts as bank conflicts).
    const int q = lane >> 3, l8 = lane & 7;
    const int lane_row = (warp >> 1) * 8 + (warp & 1) + 2 * q;
    const int n_iter = (128 + args.chunk_n + 31) >> 5;
    const uint32_t stag_lane = smem_base + kWgSmemStag + (uint32_t)l8 * 16u + (uint32_t)(lane_row * args.pitch);
    const uint32_t stag_step = (uint32_t)(32 * args.pitch);
    // tile byte offset of (row lane_row, pixels l8*4..+3 of half 0): 8 B at 16-byte chunk (l8 / 2) XOR (row & 7)
    const uint32_t tile_row = smem_base + a_row_offset(lane_row);
    const uint32_t sw = (uint32_t)lane_row & 7u;
    uint32_t ph = 0;   // parity of the next completion of stg_full[h] / tile_empty (one K-block = one phase of each)
    int it = 0;
    for (int item = blockIdx.x; item < args.n_items; item += gridDim.x, ++it) {
      int img, mt, ch, kb0, kb1;
      decode(item, img, mt, ch, kb0, kb1);
      const int m0 = mt * 128, n0 = ch * args.chunk_n;
      uint32_t ld_mask = 0u, one_mask = 0u, st_mask = 0u;
      for (int u = 0; u < n_iter; ++u) {
        const int row = u * 32 + lane_row;
        if (row < 128) {
          if (row < args.rows_dy_box && m0 + row < args.N_dy) ld_mask |= 1u << u;
          st_mask |= 1u << u;
        } else if (row < 128 + args.chunk_n) {
          const int r = row - 128, n = n0 + r;
          if (n < args.K_x && r < args.rows_x_box) ld_mask |= 1u << u;
          else if (n >= args.K_x && n < args.rows_x) one_mask |= 1u << u;   // the ones row -> bias gradient column
          st_mask |= 1u << u;
        }
      }
      if (l8 * 16 >= args.pitch) ld_mask = one_mask = 0u;   // HW < 32: this lane's pixels lie beyond the (narrower) TMA box
      for (int kb = kb0; kb < kb1; ++kb) {
        uint32_t hw[2][kWgMaxIter][2], lw[2][kWgMaxIter][2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const float one = (kb * 64 + h * 32 + l8 * 4 < args.HW) ? 1.0f : 0.0f;   // pixels past the end of the image contribute nothing
          mbar_wait(bar(kWgBarStgFull + h), ph);
#pragma unroll
          for (int u = 0; u < kWgMaxIter; ++u) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if ((ld_mask >> u) & 1u)   // inside the tensors TMA zero-fills out-of-range pixels itself
              asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                           : "r"(stag_lane + (uint32_t)(h * kWgStagBytes) + (uint32_t)u * stag_step));
            else if ((one_mask >> u) & 1u) v = make_float4(one, one, one, one);
            split2(v.x, v.y, hw[h][u][0], lw[h][u][0]);
            split2(v.z, v.w, hw[h][u][1], lw[h][u][1]);
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(bar(kWgBarStgEmpty + h));    // slot consumed: the loader refills it with the next K-block's half
        }
        mbar_wait(bar(kWgBarTileEmpty), ph ^ 1);                   // the MMAs of the previous K-block have read the operand tiles
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          // pixels h*32 + l8*4 .. +3 -> bytes h*64 + l8*8 of the 128-byte row: chunk (h*4 + l8/2) XOR sw, + (l8 & 1) * 8
          const uint32_t col = ((((uint32_t)(h * 4) + ((uint32_t)l8 >> 1)) ^ sw) << 4) + ((uint32_t)l8 & 1u) * 8u;
#pragma unroll
          for (int u = 0; u < kWgMaxIter; ++u) {
            if ((st_mask >> u) & 1u) {
              // rows advance by 32 per iteration = 4 swizzle groups = 4096 B; u >= 4 continues in the B tile
              const uint32_t base = tile_row + col + (u < 4 ? (uint32_t)(u * 4096) : (uint32_t)(kWgABytes + (u - 4) * 4096));
              const uint32_t lo_off = u < 4 ? 16384u : 32768u;
              asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(base), "r"(hw[h][u][0]), "r"(hw[h][u][1]) : "memory");
              asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(base + lo_off), "r"(lw[h][u][0]), "r"(lw[h][u][1]) : "memory");
            }
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(kWgBarTileFull));
        ph ^= 1;
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
  } else if (warp == 9) {
    // ======================================= TMA loader: two tensor loads per K-block =======================================
    if (lane == 0) {
      uint32_t ph = 0;
      const uint64_t dy_ptr = reinterpret_cast<uint64_t>(&dy_map), x_ptr = reinterpret_cast<uint64_t>(&x_map);
      for (int item = blockIdx.x; item < args.n_items; item += gridDim.x) {
        int img, mt, ch, kb0, kb1;
        decode(item, img, mt, ch, kb0, kb1);
        for (int kb = kb0; kb < kb1; ++kb) {
          for (int h = 0; h < 2; ++h) {
            mbar_wait(bar(kWgBarStgEmpty + h), ph ^ 1);
            mbar_arrive_expect_tx(bar(kWgBarStgFull + h), (uint32_t)args.stage_tx_bytes);
            const uint32_t dst = smem_base + kWgSmemStag + (uint32_t)(h * kWgStagBytes);
            asm volatile(
                "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
                "l"(dy_ptr), "r"(kb * 64 + h * 32), "r"(mt * 128), "r"(img), "r"(bar(kWgBarStgFull + h))
                : "memory");
            asm volatile(
                "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                    dst + (uint32_t)(128 * args.pitch)),
                "l"(x_ptr), "r"(kb * 64 + h * 32), "r"(ch * args.chunk_n), "r"(img), "r"(bar(kWgBarStgFull + h))
                : "memory");
          }
          ph ^= 1;
        }
      }
    }
    __syncwarp();
  } else {
    // ======================================= MMA issuer =======================================
    uint32_t ph = 0;
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t sbase_u = __shfl_sync(0xffffffffu, smem_base, 0);
    constexpr uint32_t kDescHiSw128 = (uint32_t)((1024 >> 4) | (1u << 14) | (2u << 29));
    constexpr uint32_t kDescLoLboSw = 1u << 16;
    auto mk = [](uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; };
    const uint32_t idesc = umma_idesc_bf16(128, args.chunk_n);
    const uint32_t a_hi0 = ((sbase_u >> 4) & 0x3FFFu) | kDescLoLboSw;
    const uint32_t a_lo0 = a_hi0 + (16384 >> 4);
    const uint32_t b_hi0 = a_hi0 + (kWgABytes >> 4);
    const uint32_t b_lo0 = b_hi0 + (32768 >> 4);
    int it = 0;
    for (int item = blockIdx.x; item < args.n_items; item += gridDim.x, ++it) {
      int img, mt, ch, kb0, kb1;
      decode(item, img, mt, ch, kb0, kb1);
      mbar_wait(bar(kWgBarAccEmpty), (uint32_t)((it & 1) ^ 1));
      tc_fence_after_sync();
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(bar(kWgBarTileFull), ph);
        tc_fence_after_sync();
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
          umma_commit(bar(kWgBarTileEmpty));
        }
        __syncwarp();
        ph ^= 1;
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
  int n_sm = 0;
  {
    int rc = device_once(kOnceWgradTc, &n_sm, []() -> int {
      GNRF_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgSmemBytes));
      return GNRF_OK;
    });
    if (rc != GNRF_OK) return rc;
  }
  if (HW % 4 != 0 || dy_img_stride % 4 != 0 || x_img_stride % 4 != 0)
    return fail(GNRF_ERR_ARG, "wgrad_tc: HW and image strides must be multiples of 4 floats (vector loads)");
  WgradPlan pl = wgrad_plan(N_dy, K_x, n_img, HW, db_img != nullptr);
  if (ws_bytes < pl.partial_bytes) return fail(GNRF_ERR_ARG, "wgrad_tc: workspace %zu < required %zu bytes", ws_bytes, pl.partial_bytes);
  if ((reinterpret_cast<uintptr_t>(dY) & 15) != 0 || (reinterpret_cast<uintptr_t>(X) & 15) != 0)
    return fail(GNRF_ERR_ARG, "wgrad_tc: dY and X must be 16-byte aligned (TMA)");
  TensorMapEncodeFn enc = tensor_map_encoder();
  if (enc == nullptr) return fail(GNRF_ERR_CUDA, "wgrad_tc: cuTensorMapEncodeTiled is not available from the driver");
  const cuuint32_t box_px = (cuuint32_t)(HW < 32 ? HW : 32);   // half a K-block per TMA load
  const cuuint32_t box_dy = (cuuint32_t)(N_dy < 128 ? N_dy : 128), box_x = (cuuint32_t)(K_x < pl.chunk_n ? K_x : pl.chunk_n);
  CUtensorMap dy_map, x_map;
  {
    const cuuint64_t gdim[3] = {(cuuint64_t)HW, (cuuint64_t)N_dy, (cuuint64_t)n_img};
    const cuuint64_t gstr[2] = {(cuuint64_t)HW * 4, (cuuint64_t)dy_img_stride * 4};
    const cuuint32_t box[3] = {box_px, box_dy, 1}, estr[3] = {1, 1, 1};
    CUresult cr = enc(&dy_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(dY), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return fail(GNRF_ERR_CUDA, "wgrad_tc: cuTensorMapEncodeTiled(dY) failed (%d)", (int)cr);
  }
  {
    const cuuint64_t gdim[3] = {(cuuint64_t)HW, (cuuint64_t)K_x, (cuuint64_t)n_img};
    const cuuint64_t gstr[2] = {(cuuint64_t)HW * 4, (cuuint64_t)x_img_stride * 4};
    const cuuint32_t box[3] = {box_px, box_x, 1}, estr[3] = {1, 1, 1};
    CUresult cr = enc(&x_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(X), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return fail(GNRF_ERR_CUDA, "wgrad_tc: cuTensorMapEncodeTiled(X) failed (%d)", (int)cr);
  }
  WgradArgs a;
  a.N_dy = N_dy; a.K_x = K_x; a.rows_x = pl.rows_x; a.HW = HW; a.n_img = n_img;
  a.chunk_n = pl.chunk_n; a.n_ch = pl.n_ch; a.n_mt = pl.n_mt; a.n_split = pl.n_split; a.n_kb_total = pl.n_kb_total; a.n_items = pl.n_items;
  a.pitch = (int)box_px * 4;
  a.rows_dy_box = (int)box_dy; a.rows_x_box = (int)box_x;
  a.stage_tx_bytes = (int)(box_px * 4 * (box_dy + box_x));
  a.partial = static_cast<float*>(ws);
  int grid = pl.n_items < n_sm ? pl.n_items : n_sm;
  wgrad_tc_kernel<<<grid, kWgThreads, kWgSmemBytes, st>>>(a, dy_map, x_map);
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
