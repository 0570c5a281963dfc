// Training-path dense layers on PRE-SPLIT bf16 planes (sm_100a, tcgen05 + tensor-map TMA, no conversion pass).
//
// The radiance-MLP layers of the training path (models/mlp_nerf.py:95-119, nn.Conv2d(k=1) on [B,C,N_r,N_s]; gradients: the reference
// uses torch autograd, trainer/gazenerf_trainer.py:479-528) are evaluated layer by layer on channel-major activations kept in HBM
// (gazenerf_b200/train.py).  r1 kept those activations in fp32 and every GEMM kernel re-split them into the bf16 hi/lo operand pair
// on CUDA cores on the way in (conv_tc.cu / wgrad_tc.cu: converter warps + an fp32 staging ring; 0.53 / 0.30 of the HBM roof).
// Here the PRODUCER writes the split once: an activation tensor is two bf16 planes
//     plane 0 = hi = bf16(x),   plane 1 = lo = bf16(x - hi)        each [n_img][rows][HW], points contiguous
// (the same 4 bytes per element as fp32, the same 16 significand bits the bf16x3 GEMMs consumed before), and every consumer feeds
// the planes to the tensor cores straight from TMA:
//   * lin_hl_kernel   (forward and input gradient)   D^T[out ch][pt] = W[out ch][k] . X[k][pt]
//       UMMA M = 128 output channels (TMEM lane = channel), N = 256 points, K = 32 channels per stage.
//       A = packed weights (K-major no-swizzle core matrices, bulk copy), B = activation planes, which are MN-major for this
//       product (points contiguous): SWIZZLE_128B MN-major atoms [64 pt x 8 ch] written by tensor-map TMA (one box per 64-point group).
//       Two 256-column TMEM accumulators: the epilogue of (tile, M-tile) i overlaps the MMAs of i+1.  Epilogue thread = output
//       channel, so it owns 32 CONSECUTIVE points per tcgen05.ld: bias (+per-image bias), ReLU or the saved ReLU sign bits (input
//       gradient; the forward writes them as a 1-bit-per-element tensor), hi/lo split, then a swizzled smem slot per warp and ONE TMA
//       tensor store of [planes][32 rows][32 points] (direct per-thread stores touched 32 different lines per instruction and held the
//       kernel at 2 TB/s of output); fp32 rows for the non-GEMM consumers go out as plain vector stores.
//   * wgrad_hl_kernel (weight gradient)               dW[out][in] = sum_pt dY[out][pt] X[in][pt]
//       both operands K-major (K = points) SWIZZLE_64B tiles of 32 points straight from TMA; M = 128 rows of dY, N = ALL rows of X
//       (<= 392, two UMMA N-blocks) so that dY is read once and X n_mt times (r1: 2x / 3x); an all-ones row group in front of X
//       gives the bias gradient in accumulator column 0; deterministic split-K (partials + fixed-order reduce).
// PL = 2 is the bf16x3 scheme (x*w ~= hi*hi + lo*hi + hi*lo, fp32 accumulate), PL = 1 reads / writes the hi plane only: plain
// single-pass bf16 (BASELINE config[4] names bf16), half the bytes, a third of the MMAs, its own stated tolerance in the tests.
#include <cuda.h>   // CUtensorMap (types only; the encoder comes through cudaGetDriverEntryPoint)

#include "common.cuh"
#include "tc_common.cuh"

namespace gnrf {
namespace tc {

// ===================================================================================================== lin_hl (forward / dgrad)
constexpr int kLhTileN = 256;                    // points per tile (UMMA N)
constexpr int kLhKb = 32;                        // channels per pipeline stage (two K16 steps)
constexpr int kLhThreads = 320;                  // warps 0-7 epilogue, 8 loader, 9 MMA
constexpr int kLhActPlane = 4 * kLhKb * 128;     // 4 point groups x [32 ch][64 pt] bf16 = 16 KB per plane (stage: group-major, planes inside)
constexpr int kLhWPlane = 128 * kLhKb * 2;       // [2 K16 slices][128 rows x 32 B] = 8 KB

template <int PL>
struct LhCfg {
  static constexpr int kStageBytes = PL * (kLhActPlane + kLhWPlane);
  static constexpr int kStages = PL == 2 ? 4 : 7;   // bytes in flight per SM bound the L2 -> SM stream (4 x 48 KB measured 5 % faster
  static constexpr int kSlots = PL == 2 ? 1 : 2;    // than 3 x 48 KB + a second output staging slot per epilogue warp)
  static constexpr int kWOff = PL * kLhActPlane;
  static constexpr int kBarFull = 0, kBarEmpty = kStages, kBarAccFull = 2 * kStages, kBarAccEmpty = 2 * kStages + 2,
                       kNumBars = 2 * kStages + 4;
  // output staging: per epilogue warp kSlots slots of [PL planes][32 rows][64 B] (SWIZZLE_64B), drained by TMA tensor stores
  static constexpr int kSlotBytes = PL * 2048;
  static constexpr int kSmemStage = kStages * kStageBytes;
  static constexpr int kSmemBars = kSmemStage + 8 * kSlots * kSlotBytes;
  static constexpr int kSmemMisc = kSmemBars + kNumBars * 8;
  static constexpr int kSmemBytes = kSmemMisc + 64 + 1024;
};

struct LinArgs {
  const unsigned char* wpack;   // [n_mt][n_kb][PL planes][2 K16 slices][128 rows x 32 B]
  const float* bias;            // [n_mt * 128] (zero padded), follows the blobs in the packed buffer
  const float* bias_img;        // [n_img][N_out] or null
  int N_out, K_in, n_mt, n_kb, HW, n_img, tiles_per_img, n_tiles, act, act_rows;   // ReLU on rows < act_rows (a multiple of 32, or >= N_out)
  __nv_bfloat16* out;           // planes: rows < hl_rows
  long long out_img_stride, out_plane_stride;
  int hl_rows;
  float* out_f32;               // fp32 rows: output row r >= hl_rows -> out_f32[img][r - hl_rows][pt]
  long long f32_img_stride;
  const uint32_t* mask_bits;    // [n_img][rows][HW/32] sign bits of a saved post-ReLU activation: v = bit ? v : 0 for rows < mask_rows
  long long mask_img_stride;    // in 32-bit words
  int mask_rows;
  uint32_t* mask_out;           // nullable: bit j of word [img][row][p/32] = (output > 0), rows < mask_out_rows
  long long mask_out_img_stride;
  int mask_out_rows;
};

__device__ __forceinline__ void st_global_v4(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

template <int PL>
__global__ void __launch_bounds__(kLhThreads, 1) lin_hl_kernel(const LinArgs args, const __grid_constant__ CUtensorMap x_map,
                                                               const __grid_constant__ CUtensorMap out_map) {
  using C = LhCfg<PL>;
  extern __shared__ unsigned char smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  unsigned char* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bars = smem_base + C::kSmemBars;
  auto bar = [&](int i) { return bars + (uint32_t)i * 8u; };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem_gen + C::kSmemMisc);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < C::kStages; ++i) { mbar_init(bar(C::kBarFull + i), 1); mbar_init(bar(C::kBarEmpty + i), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(bar(C::kBarAccFull + i), 1); mbar_init(bar(C::kBarAccEmpty + i), 8); }
    fence_mbar_init();
  }
  if (warp == 9) tmem_alloc_512(smem_u32(const_cast<uint32_t*>(tmem_ptr_smem)));
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr_smem;
  // item = (tile, M-tile), M-tile fastest ACROSS CTAs: the n_mt CTAs that need the same activation tile read it at the same time, so it
  // comes from HBM once (with one CTA walking the M-tiles of its tile, the re-reads 10 and 20 us later missed L2: 1.5x DRAM reads).
  // Tried and rejected: those n_mt CTAs as a cluster with the tile multicast (every SM still ingests the same bytes -- the limit is
  // the per-SM L2 -> SM stream -- and the lockstep cost 15 %); L2 prefetch of the next tile (no gain here, 2x DRAM reads in wgrad_hl).
  const int n_items = args.n_tiles * args.n_mt;

  if (warp < 8) {
    // ======================================= epilogue: thread = output channel, 4 x 32 consecutive points =======================
    const int q = warp & 3, hf = warp >> 2;
    const int words_per_row = args.HW >> 5;
    const uint64_t out_map_ptr = reinterpret_cast<uint64_t>(&out_map);
    const uint32_t slot0 = smem_base + (uint32_t)C::kSmemStage + (uint32_t)(warp * C::kSlots * C::kSlotBytes);
    const uint32_t my_off = (uint32_t)(lane * 64), sw = ((uint32_t)lane >> 1) & 3u;
    uint32_t n_st = 0;   // TMA stores issued by this warp (slot = n_st & 1)
    int it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int tile = item / args.n_mt, mt = item - tile * args.n_mt;
      const int img = tile / args.tiles_per_img, p_tile = (tile - img * args.tiles_per_img) * kLhTileN + hf * 128;
      {
        const int buf = it & 1;
        const int row_base = mt * 128 + q * 32, row = row_base + lane;
        const bool warp_live = row_base < args.N_out, warp_planes = row_base < args.hl_rows;   // hl_rows % 32 == 0 or >= N_out
        const bool relu = args.act != 0 && row_base < args.act_rows;
        const bool valid = row < args.N_out;
        float b = 0.0f;
        if (valid) {
          b = args.bias[row];
          if (args.bias_img != nullptr) b += args.bias_img[(size_t)img * args.N_out + row];
        }
        const bool use_mask = valid && args.mask_bits != nullptr && row < args.mask_rows;
        uint4 mw4 = make_uint4(~0u, ~0u, ~0u, ~0u);
        if (use_mask)
          mw4 = *reinterpret_cast<const uint4*>(args.mask_bits + (size_t)img * args.mask_img_stride + (size_t)row * words_per_row + (p_tile >> 5));
        const uint32_t mw[4] = {mw4.x, mw4.y, mw4.z, mw4.w};
        uint32_t mo[4] = {0u, 0u, 0u, 0u};
        mbar_wait(bar(C::kBarAccFull + buf), (uint32_t)((it >> 1) & 1));
        tc_fence_after_sync();
        const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * 256 + hf * 128);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t r[32];
          tmem_ld32(t_addr + (uint32_t)(c * 32), r);
          tmem_wait_ld();
          if (warp_live) {
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              v[j] = __uint_as_float(r[j]) + b;
              if (relu) v[j] = fmaxf(v[j], 0.0f);
            }
            if (args.mask_bits != nullptr) {   // warp-uniform: the input-gradient calls
              const uint32_t m = mw[c];
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = ((m >> j) & 1u) ? v[j] : 0.0f;
            }
            if (args.mask_out != nullptr) {    // warp-uniform: the forward calls
#pragma unroll
              for (int j = 0; j < 32; ++j) mo[c] |= (v[j] > 0.0f ? 1u : 0u) << j;
            }
            const int p0 = p_tile + c * 32;
            if (warp_planes) {
              uint32_t hi[16], lo[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) split2(v[2 * j], v[2 * j + 1], hi[j], lo[j]);
              if (n_st >= (uint32_t)C::kSlots) {   // the store that last used this slot has finished READING it
                if (lane == 0) {
                  if (C::kSlots == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                  else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                }
                __syncwarp();
              }
              const uint32_t slot = slot0 + (n_st % (uint32_t)C::kSlots) * (uint32_t)C::kSlotBytes;
#pragma unroll
              for (int j = 0; j < 4; ++j)
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(slot + my_off + (((uint32_t)j ^ sw) << 4)), "r"(hi[4 * j]),
                             "r"(hi[4 * j + 1]), "r"(hi[4 * j + 2]), "r"(hi[4 * j + 3])
                             : "memory");
              if (PL == 2) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(slot + 2048u + my_off + (((uint32_t)j ^ sw) << 4)),
                               "r"(lo[4 * j]), "r"(lo[4 * j + 1]), "r"(lo[4 * j + 2]), "r"(lo[4 * j + 3])
                               : "memory");
              }
              fence_proxy_async_smem();
              __syncwarp();
              if (lane == 0) {   // one tensor store: [PL planes][32 rows][32 points]; rows past the tensor are clipped by the TMA unit
                asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3, %4}], [%5];" ::"l"(out_map_ptr),
                             "r"(p0), "r"(row_base), "r"(img), "r"(0), "r"(slot)
                             : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
              }
              ++n_st;
            } else if (valid) {
              float* o = args.out_f32 + (size_t)img * args.f32_img_stride + (size_t)(row - args.hl_rows) * args.HW + p0;
#pragma unroll
              for (int j = 0; j < 8; ++j)
                st_global_v4(o + j * 4, __float_as_uint(v[4 * j]), __float_as_uint(v[4 * j + 1]), __float_as_uint(v[4 * j + 2]),
                             __float_as_uint(v[4 * j + 3]));
            }
          }
        }
        if (valid && args.mask_out != nullptr && row < args.mask_out_rows)
          *reinterpret_cast<uint4*>(args.mask_out + (size_t)img * args.mask_out_img_stride + (size_t)row * words_per_row + (p_tile >> 5)) =
              make_uint4(mo[0], mo[1], mo[2], mo[3]);
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(C::kBarAccEmpty + buf));
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // all output stores complete before the CTA exits
    __syncwarp();
  } else if (warp == 8) {
    // ======================================= loader: one tensor-map TMA (activations) + one bulk copy (weights) per stage =======
    if (elect_one()) {
      const uint64_t map_ptr = reinterpret_cast<uint64_t>(&x_map);
      uint32_t s = 0, ph = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int tile = item / args.n_mt, mt = item - tile * args.n_mt;
        const int img = tile / args.tiles_per_img, p0 = (tile - img * args.tiles_per_img) * kLhTileN;
        {
          const unsigned char* wsrc = args.wpack + (size_t)mt * args.n_kb * (PL * kLhWPlane);
          for (int kb = 0; kb < args.n_kb; ++kb) {
            mbar_wait_spin(bar(C::kBarEmpty + s), ph ^ 1);
            const uint32_t dst = smem_base + s * (uint32_t)C::kStageBytes, fb = bar(C::kBarFull + s);
            mbar_arrive_expect_tx(fb, (uint32_t)C::kStageBytes);
#pragma unroll
            for (int g = 0; g < 4; ++g)   // one [PL planes][32 ch][64 pt] box per 64-point group: group g = [hi 4 KB | lo 4 KB]
              asm volatile(
                  "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::
                      "r"(dst + (uint32_t)(g * PL * 4096)), "l"(map_ptr), "r"(p0 + g * 64), "r"(kb * kLhKb), "r"(img), "r"(0), "r"(fb)
                  : "memory");
            bulk_g2s(dst + (uint32_t)C::kWOff, wsrc + (size_t)kb * (PL * kLhWPlane), (uint32_t)(PL * kLhWPlane), fb);
            if (++s == (uint32_t)C::kStages) { s = 0; ph ^= 1; }
          }
        }
      }
    }
    __syncwarp();
  } else {
    // ======================================= MMA issuer (one elected thread runs the whole loop) ================================
    if (elect_one()) {
      // B = activations, MN-major SWIZZLE_128B: 64-point groups LBO = PL * 4096 B apart (group = [hi 4 KB | lo 4 KB]), 8-channel
      // groups SBO = 1024 B apart
      constexpr uint32_t kBHi = (uint32_t)((1024 >> 4) | (1u << 14) | (2u << 29));
      constexpr uint32_t kBLoFlags = (uint32_t)((PL * 4096) >> 4) << 16;
      // A = weights, K-major no-swizzle core matrices: K halves LBO = 128 B, 8-row groups SBO = 256 B
      constexpr uint32_t kAHi = (uint32_t)((256 >> 4) | (1u << 14));
      constexpr uint32_t kALoFlags = (uint32_t)(128 >> 4) << 16;
      auto mk = [](uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; };
      const uint32_t idesc = umma_idesc_bf16(128, kLhTileN) | (1u << 16);   // bit 16: B is MN-major
      uint32_t s = 0, ph = 0;
      int it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        {
          const int buf = it & 1;
          mbar_wait_spin(bar(C::kBarAccEmpty + buf), (uint32_t)(((it >> 1) & 1) ^ 1));
          tc_fence_after_sync();
          const uint32_t d = tmem_base + (uint32_t)(buf * 256);
          for (int kb = 0; kb < args.n_kb; ++kb) {
            mbar_wait_spin(bar(C::kBarFull + s), ph);
            tc_fence_after_sync();
            const uint32_t sb = smem_base + s * (uint32_t)C::kStageBytes;
#pragma unroll
            for (int k16 = 0; k16 < 2; ++k16) {
              const uint64_t b_hi = mk((((sb + (uint32_t)(k16 * 2048)) >> 4) & 0x3FFFu) | kBLoFlags, kBHi);
              const uint64_t a_hi = mk((((sb + (uint32_t)(C::kWOff + k16 * 4096)) >> 4) & 0x3FFFu) | kALoFlags, kAHi);
              if (PL == 2) {
                // W_hi feeds two consecutive UMMAs from the tensor core's A collector (one smem read of the 4 KB slice instead of two)
                const uint64_t b_lo = mk((((sb + (uint32_t)(4096 + k16 * 2048)) >> 4) & 0x3FFFu) | kBLoFlags, kBHi);
                const uint64_t a_lo = mk((((sb + (uint32_t)(C::kWOff + kLhWPlane + k16 * 4096)) >> 4) & 0x3FFFu) | kALoFlags, kAHi);
                umma_ss_a_fill(d, a_hi, b_hi, idesc, (kb == 0 && k16 == 0) ? 0u : 1u);
                umma_ss_a_lastuse(d, a_hi, b_lo, idesc, 1u);
                umma_ss(d, a_lo, b_hi, idesc, 1u);
              } else {
                umma_ss(d, a_hi, b_hi, idesc, (kb == 0 && k16 == 0) ? 0u : 1u);
              }
            }
            umma_commit(bar(C::kBarEmpty + s));
            if (++s == (uint32_t)C::kStages) { s = 0; ph ^= 1; }
          }
          umma_commit(bar(C::kBarAccFull + buf));
        }
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

// W(n,k) = W[n * sn + k * sk] -> blobs [mt][kb][plane][K16 slice][row group][K half][row in group][8 bf16], then the bias.
__global__ void lin_hl_pack_kernel(const float* __restrict__ W, long long sn, long long sk, const float* __restrict__ b, int N, int K,
                                   int n_mt, int n_kb, int planes, unsigned char* __restrict__ dst, float* __restrict__ bias_out) {
  const size_t units_per_blob = 2 * 128 * 2;   // 16-byte units of one plane of one (mt, kb) blob
  const size_t total = (size_t)n_mt * n_kb * units_per_blob;
  for (size_t u = (size_t)blockIdx.x * blockDim.x + threadIdx.x; u < total; u += (size_t)gridDim.x * blockDim.x) {
    const size_t blob = u / units_per_blob;
    const int rem = (int)(u % units_per_blob);
    const int k16 = rem >> 8, in = rem & 255;
    const int row = (in >> 4) * 8 + (in & 7), k_half = (in >> 3) & 1;
    const int mt = (int)(blob / n_kb), kb = (int)(blob % n_kb);
    const int n = mt * 128 + row, k0 = kb * kLhKb + k16 * 16 + k_half * 8;
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int qd = 0; qd < 4; ++qd) {
      const int ka = k0 + 2 * qd, kc = ka + 1;
      const float a = (n < N && ka < K) ? W[(size_t)n * sn + (size_t)ka * sk] : 0.0f;
      const float c = (n < N && kc < K) ? W[(size_t)n * sn + (size_t)kc * sk] : 0.0f;
      split2(a, c, hi[qd], lo[qd]);
    }
    unsigned char* o = dst + blob * (size_t)(planes * kLhWPlane) + (size_t)rem * 16;
    *reinterpret_cast<uint4*>(o) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    if (planes == 2) *reinterpret_cast<uint4*>(o + kLhWPlane) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_mt * 128; i += gridDim.x * blockDim.x)
    bias_out[i] = (i < N && b != nullptr) ? b[i] : 0.0f;
}

typedef CUresult (*TensorMapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                      const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static TensorMapEncodeFn hl_tensor_map_encoder() {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
    return reinterpret_cast<TensorMapEncodeFn>(p);
  return nullptr;
}

struct LinPlan {
  int n_mt, n_kb;
  size_t blob_bytes, total_bytes;
};
static LinPlan lin_plan(int N, int K, int planes) {
  LinPlan pl;
  pl.n_mt = (N + 127) / 128;
  pl.n_kb = (K + kLhKb - 1) / kLhKb;
  pl.blob_bytes = (size_t)pl.n_mt * pl.n_kb * planes * kLhWPlane;
  pl.total_bytes = pl.blob_bytes + (size_t)pl.n_mt * 128 * sizeof(float);
  return pl;
}

template <int PL>
static int lin_hl_launch_t(const LinArgs& a, const CUtensorMap& map, const CUtensorMap& out_map, int n_sm, cudaStream_t st) {
  int dummy = 0;
  int rc = device_once(PL == 2 ? kOnceLinHl2 : kOnceLinHl1, &dummy, []() -> int {
    GNRF_CUDA(cudaFuncSetAttribute(lin_hl_kernel<PL>, cudaFuncAttributeMaxDynamicSharedMemorySize, LhCfg<PL>::kSmemBytes));
    return GNRF_OK;
  });
  if (rc != GNRF_OK) return rc;
  const int n_items = a.n_tiles * a.n_mt;
  const int grid = n_items < n_sm ? n_items : n_sm;
  lin_hl_kernel<PL><<<grid, kLhThreads, LhCfg<PL>::kSmemBytes, st>>>(a, map, out_map);
  return GNRF_OK;
}

// ===================================================================================================== wgrad_hl (weight gradient)
constexpr int kWhKb = 32;                        // points per stage (two K16 steps), SWIZZLE_64B rows of 64 B
constexpr int kWhThreads = 192;                  // warps 0-3 drain, 4 loader, 5 MMA
constexpr int kWhAPlane = 128 * 64;              // 8 KB
constexpr int kWhBRowsMax = 400;                 // 8 (ones group) + <= 384 X rows + pad to a multiple of 16
constexpr int kWhBPlane = kWhBRowsMax * 64;      // 25 600 B

template <int PL>
struct WhCfg {
  static constexpr int kStageBytes = PL * (kWhAPlane + kWhBPlane);
  static constexpr int kStages = PL == 2 ? 3 : 6;
  static constexpr int kBOff = PL * kWhAPlane;
  static constexpr int kBarFull = 0, kBarEmpty = kStages, kBarAccFull = 2 * kStages, kBarAccEmpty = 2 * kStages + 1,
                       kNumBars = 2 * kStages + 2;
  static constexpr int kSmemBars = kStages * kStageBytes;
  static constexpr int kSmemMisc = kSmemBars + kNumBars * 8;
  static constexpr int kSmemBytes = kSmemMisc + 64 + 1024;
};

struct WgHlArgs {
  int N_dy, K_x, n_img;
  int n_mt, n_ch, chunk_rows, chunk_n, n_box, box_rows, n_split, n_kb_total, n_items;
  int nb0, nb1;               // UMMA N of the two column blocks (nb1 = 0: one block)
  int stage_tx_bytes;
  float* partial;             // [n_items][128][chunk_n]
};

template <int PL>
__global__ void __launch_bounds__(kWhThreads, 1) wgrad_hl_kernel(const WgHlArgs args, const __grid_constant__ CUtensorMap dy_map,
                                                                 const __grid_constant__ CUtensorMap x_map) {
  using C = WhCfg<PL>;
  extern __shared__ unsigned char smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  unsigned char* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bars = smem_base + C::kSmemBars;
  auto bar = [&](int i) { return bars + (uint32_t)i * 8u; };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem_gen + C::kSmemMisc);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // B planes: zero everything once (pad rows stay zero: the TMA boxes cover exactly the X rows), then the ones group of the hi plane
  for (int s = 0; s < C::kStages; ++s) {
    uint4* bz = reinterpret_cast<uint4*>(smem_gen + (size_t)s * C::kStageBytes + C::kBOff);
    for (int i = threadIdx.x; i < PL * kWhBPlane / 16; i += kWhThreads) bz[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  __syncthreads();
  for (int s = 0; s < C::kStages; ++s) {
    uint32_t* one = reinterpret_cast<uint32_t*>(smem_gen + (size_t)s * C::kStageBytes + C::kBOff);
    for (int i = threadIdx.x; i < 512 / 4; i += kWhThreads) one[i] = 0x3F803F80u;   // bf16 1.0 pairs (swizzle-invariant)
  }
  fence_proxy_async_smem();
  if (threadIdx.x == 0) {
    for (int i = 0; i < C::kStages; ++i) { mbar_init(bar(C::kBarFull + i), 1); mbar_init(bar(C::kBarEmpty + i), 1); }
    mbar_init(bar(C::kBarAccFull), 1);
    mbar_init(bar(C::kBarAccEmpty), 4);
    fence_mbar_init();
  }
  if (warp == 5) tmem_alloc_512(smem_u32(const_cast<uint32_t*>(tmem_ptr_smem)));
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

  if (warp < 4) {
    // ======================================= drain: thread = dY row, 32 columns at a time -> partial[item][row][col] =============
    int it = 0;
    for (int item = blockIdx.x; item < args.n_items; item += gridDim.x, ++it) {
      mbar_wait(bar(C::kBarAccFull), (uint32_t)(it & 1));
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
          if (j < ncols) st_global_v4(dst + c0 + j, r[j], r[j + 1], r[j + 2], r[j + 3]);
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(C::kBarAccEmpty));
    }
  } else if (warp == 4) {
    // ======================================= loader =================================================================================
    if (elect_one()) {
      const uint64_t dy_ptr = reinterpret_cast<uint64_t>(&dy_map), x_ptr = reinterpret_cast<uint64_t>(&x_map);
      uint32_t s = 0, ph = 0;
      for (int item = blockIdx.x; item < args.n_items; item += gridDim.x) {
        int img, mt, ch, kb0, kb1;
        decode(item, img, mt, ch, kb0, kb1);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait_spin(bar(C::kBarEmpty + s), ph ^ 1);
          const uint32_t dst = smem_base + s * (uint32_t)C::kStageBytes, fb = bar(C::kBarFull + s);
          mbar_arrive_expect_tx(fb, (uint32_t)args.stage_tx_bytes);
          asm volatile(
              "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(dst),
              "l"(dy_ptr), "r"(kb * kWhKb), "r"(mt * 128), "r"(img), "r"(0), "r"(fb)
              : "memory");
          for (int pl = 0; pl < PL; ++pl)
            for (int bx = 0; bx < args.n_box; ++bx)
              asm volatile(
                  "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(
                      dst + (uint32_t)(C::kBOff + pl * kWhBPlane + 512 + bx * args.box_rows * 64)),
                  "l"(x_ptr), "r"(kb * kWhKb), "r"(ch * args.chunk_rows + bx * args.box_rows), "r"(img), "r"(pl), "r"(fb)
                  : "memory");
          if (++s == (uint32_t)C::kStages) { s = 0; ph ^= 1; }
        }
      }
    }
    __syncwarp();
  } else {
    // ======================================= MMA issuer =============================================================================
    if (elect_one()) {
      // K-major SWIZZLE_64B: rows of 64 B (32 points), 8-row groups SBO = 512 B apart
      constexpr uint32_t kDescHi = (uint32_t)((512 >> 4) | (1u << 14) | (4u << 29));
      constexpr uint32_t kDescLoFlags = 1u << 16;
      auto mk = [](uint32_t addr) { return ((uint64_t)kDescHi << 32) | (((addr >> 4) & 0x3FFFu) | kDescLoFlags); };
      const uint32_t idesc0 = umma_idesc_bf16(128, args.nb0), idesc1 = umma_idesc_bf16(128, args.nb1 > 0 ? args.nb1 : 16);
      uint32_t s = 0, ph = 0;
      int it = 0;
      for (int item = blockIdx.x; item < args.n_items; item += gridDim.x, ++it) {
        int img, mt, ch, kb0, kb1;
        decode(item, img, mt, ch, kb0, kb1);
        mbar_wait_spin(bar(C::kBarAccEmpty), (uint32_t)((it & 1) ^ 1));
        tc_fence_after_sync();
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait_spin(bar(C::kBarFull + s), ph);
          tc_fence_after_sync();
          const uint32_t sa = smem_base + s * (uint32_t)C::kStageBytes, sbm = sa + (uint32_t)C::kBOff;
#pragma unroll
          for (int k16 = 0; k16 < 2; ++k16) {
            const uint32_t acc = (kb == kb0 && k16 == 0) ? 0u : 1u;
            const uint64_t a_hi = mk(sa + (uint32_t)(k16 * 32)), a_lo = mk(sa + (uint32_t)(kWhAPlane + k16 * 32));
            {
              const uint64_t b_hi = mk(sbm + (uint32_t)(k16 * 32)), b_lo = mk(sbm + (uint32_t)(kWhBPlane + k16 * 32));
              umma_ss(tmem_base, a_hi, b_hi, idesc0, acc);
              if (PL == 2) {
                umma_ss(tmem_base, a_lo, b_hi, idesc0, 1u);
                umma_ss(tmem_base, a_hi, b_lo, idesc0, 1u);
              }
            }
            if (args.nb1 > 0) {
              const uint32_t off = (uint32_t)(args.nb0 * 64);
              const uint64_t b_hi = mk(sbm + off + (uint32_t)(k16 * 32)), b_lo = mk(sbm + off + (uint32_t)(kWhBPlane + k16 * 32));
              umma_ss(tmem_base + (uint32_t)args.nb0, a_hi, b_hi, idesc1, acc);
              if (PL == 2) {
                umma_ss(tmem_base + (uint32_t)args.nb0, a_lo, b_hi, idesc1, 1u);
                umma_ss(tmem_base + (uint32_t)args.nb0, a_hi, b_lo, idesc1, 1u);
              }
            }
          }
          umma_commit(bar(C::kBarEmpty + s));
          if (++s == (uint32_t)C::kStages) { s = 0; ph ^= 1; }
        }
        umma_commit(bar(C::kBarAccFull));
      }
    }
    __syncwarp();
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after_sync();
    tmem_dealloc_512(tmem_base);
  }
}

// dW[n][k] (=|+=) sum over images and splits of partial[..][n][8 + k - ch*chunk_rows]; column 0 of chunk 0 = bias gradient.
__global__ void wgrad_hl_reduce_kernel(const float* __restrict__ partial, int N_dy, int K_x, int n_img, int n_mt, int n_ch, int n_split,
                                       int chunk_rows, int chunk_n, float* __restrict__ dW, float* __restrict__ db, int db_sum) {
  const long long total = (long long)N_dy * (K_x + 1);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i / (K_x + 1)), kk = (int)(i % (K_x + 1));   // kk == K_x: the bias column
    const int mt = n >> 7, m = n & 127;
    const int ch = kk == K_x ? 0 : kk / chunk_rows;
    const int c = kk == K_x ? 0 : 8 + kk - ch * chunk_rows;
    if (kk == K_x && db == nullptr) continue;
    float tot = 0.0f;
    for (int img = 0; img < n_img; ++img) {
      float s = 0.0f;
      const size_t item0 = (((size_t)img * n_mt + mt) * n_ch + ch) * n_split;
      for (int sp = 0; sp < n_split; ++sp) s += partial[((item0 + sp) * 128 + m) * chunk_n + c];
      if (kk == K_x && !db_sum) db[(size_t)img * N_dy + n] = s;
      tot += s;
    }
    if (kk == K_x) {
      if (db_sum) db[n] = tot;
    } else {
      dW[(size_t)n * K_x + kk] = tot;
    }
  }
}

struct WgHlPlan {
  int n_mt, n_ch, chunk_rows, chunk_n, n_box, box_rows, n_kb_total, n_split, n_items, nb0, nb1;
  size_t partial_bytes;
};
static WgHlPlan wgrad_hl_plan(int N_dy, int K_x, int n_img, int HW) {
  WgHlPlan pl;
  pl.n_mt = (N_dy + 127) / 128;
  pl.n_ch = (K_x + 383) / 384;
  int per = (K_x + pl.n_ch - 1) / pl.n_ch;
  pl.chunk_rows = ((per + 15) / 16) * 16;
  pl.n_box = pl.chunk_rows > 256 ? 2 : 1;
  pl.box_rows = pl.chunk_rows / pl.n_box;
  if (pl.n_box == 1 && pl.box_rows > K_x) pl.box_rows = K_x;   // a box never exceeds the tensor; rows past it stay zero in smem
  pl.chunk_n = ((8 + pl.chunk_rows + 15) / 16) * 16;
  if (pl.chunk_n > 256) { pl.nb0 = 208; pl.nb1 = pl.chunk_n - 208; } else { pl.nb0 = pl.chunk_n; pl.nb1 = 0; }
  pl.n_kb_total = HW / kWhKb;
  int groups = n_img * pl.n_mt * pl.n_ch;
  int sp = 148 / (groups > 0 ? groups : 1);
  if (sp < 1) sp = 1;
  if (sp > pl.n_kb_total) sp = pl.n_kb_total;
  pl.n_split = sp;
  pl.n_items = groups * sp;
  pl.partial_bytes = (size_t)pl.n_items * 128 * pl.chunk_n * sizeof(float);
  return pl;
}

template <int PL>
static int wgrad_hl_launch_t(const WgHlArgs& a, const CUtensorMap& dy_map, const CUtensorMap& x_map, int n_sm, cudaStream_t st) {
  int dummy = 0;
  int rc = device_once(PL == 2 ? kOnceWgHl2 : kOnceWgHl1, &dummy, []() -> int {
    GNRF_CUDA(cudaFuncSetAttribute(wgrad_hl_kernel<PL>, cudaFuncAttributeMaxDynamicSharedMemorySize, WhCfg<PL>::kSmemBytes));
    return GNRF_OK;
  });
  if (rc != GNRF_OK) return rc;
  const int grid = a.n_items < n_sm ? a.n_items : n_sm;
  wgrad_hl_kernel<PL><<<grid, kWhThreads, WhCfg<PL>::kSmemBytes, st>>>(a, dy_map, x_map);
  return GNRF_OK;
}

}  // namespace tc
}  // namespace gnrf

using namespace gnrf;

extern "C" size_t gnrf_lin_hl_packed_bytes(int N, int K, int planes) {
  if (N <= 0 || K <= 0 || (planes != 1 && planes != 2)) return 0;
  return (tc::lin_plan(N, K, planes).total_bytes + 255) & ~(size_t)255;
}

extern "C" int gnrf_lin_hl_pack(const float* W, const float* bias, int N, int K, int transposed, int planes, void* packed,
                                gnrf_stream_t stream) {
  GNRF_CHECK_ARG(W && packed && N > 0 && K > 0 && (planes == 1 || planes == 2));
  GNRF_CHECK_ARG((reinterpret_cast<uintptr_t>(packed) & 127) == 0);
  const tc::LinPlan pl = tc::lin_plan(N, K, planes);
  unsigned char* dst = static_cast<unsigned char*>(packed);
  const size_t units = (size_t)pl.n_mt * pl.n_kb * 512;
  tc::lin_hl_pack_kernel<<<(unsigned)((units + 255) / 256), 256, 0, as_stream(stream)>>>(
      W, transposed ? 1 : K, transposed ? N : 1, bias, N, K, pl.n_mt, pl.n_kb, planes, dst, reinterpret_cast<float*>(dst + pl.blob_bytes));
  GNRF_LAUNCH_CHECK();
  count_launches(1);
  return GNRF_OK;
}

extern "C" int gnrf_lin_hl(const void* packed, int N, int K, int planes, const void* X, long long x_img_stride, long long x_plane_stride,
                           const float* bias_img, int act, int act_rows, void* out, long long out_img_stride, long long out_plane_stride,
                           int hl_rows,
                           float* out_f32, long long f32_img_stride, const void* mask_bits, long long mask_img_stride, int mask_rows,
                           void* mask_out, long long mask_out_img_stride, int mask_out_rows, int n_img, int HW, gnrf_stream_t stream) {
  GNRF_CHECK_ARG(packed && X && N > 0 && K > 0 && n_img > 0 && HW > 0 && (planes == 1 || planes == 2));
  GNRF_CHECK_ARG(HW % tc::kLhTileN == 0 && K >= tc::kLhKb);
  GNRF_CHECK_ARG(hl_rows >= 0 && (hl_rows == 0 || out != nullptr) && (hl_rows >= N || out_f32 != nullptr));
  GNRF_CHECK_ARG(hl_rows >= N || hl_rows % 32 == 0);   // a 32-row epilogue group is either all planes or all fp32
  GNRF_CHECK_ARG(act_rows >= 0 && (act_rows >= N || act_rows % 32 == 0));
  GNRF_CHECK_ARG((reinterpret_cast<uintptr_t>(X) & 15) == 0 && x_img_stride % 8 == 0 && x_plane_stride % 8 == 0);
  GNRF_CHECK_ARG((reinterpret_cast<uintptr_t>(out) & 15) == 0 && (reinterpret_cast<uintptr_t>(out_f32) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(mask_bits) & 15) == 0 && (reinterpret_cast<uintptr_t>(mask_out) & 15) == 0 &&
                 out_img_stride % 8 == 0 && out_plane_stride % 8 == 0 && f32_img_stride % 4 == 0 && mask_img_stride % 4 == 0 &&
                 mask_out_img_stride % 4 == 0);
  int n_sm = 0;
  {
    int rc = device_once(kOnceLinHlSm, &n_sm, []() -> int { return GNRF_OK; });
    if (rc != GNRF_OK) return rc;
  }
  const tc::LinPlan pl = tc::lin_plan(N, K, planes);
  tc::TensorMapEncodeFn enc = tc::hl_tensor_map_encoder();
  if (enc == nullptr) return fail(GNRF_ERR_CUDA, "lin_hl: cuTensorMapEncodeTiled is not available from the driver");
  if (x_img_stride <= 0) x_img_stride = (long long)K * HW;
  if (planes == 2 && x_plane_stride <= 0) return fail(GNRF_ERR_ARG, "lin_hl: x_plane_stride required with 2 planes");
  CUtensorMap map;
  {
    const cuuint64_t gdim[4] = {(cuuint64_t)HW, (cuuint64_t)K, (cuuint64_t)n_img, (cuuint64_t)planes};
    const cuuint64_t gstr[3] = {(cuuint64_t)HW * 2, (cuuint64_t)x_img_stride * 2,
                                (cuuint64_t)(planes == 2 ? x_plane_stride : x_img_stride * n_img) * 2};
    const cuuint32_t box[4] = {64, (cuuint32_t)tc::kLhKb, 1, (cuuint32_t)planes}, estr[4] = {1, 1, 1, 1};
    CUresult cr = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(X), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return fail(GNRF_ERR_CUDA, "lin_hl: cuTensorMapEncodeTiled failed (%d)", (int)cr);
  }
  tc::LinArgs a;
  a.wpack = static_cast<const unsigned char*>(packed);
  a.bias = reinterpret_cast<const float*>(a.wpack + pl.blob_bytes);
  a.bias_img = bias_img;
  a.N_out = N; a.K_in = K; a.n_mt = pl.n_mt; a.n_kb = pl.n_kb; a.HW = HW; a.n_img = n_img;
  a.tiles_per_img = HW / tc::kLhTileN; a.n_tiles = a.tiles_per_img * n_img; a.act = act;
  a.act_rows = act_rows > 0 ? act_rows : N;
  a.out = static_cast<__nv_bfloat16*>(out);
  a.out_img_stride = out_img_stride > 0 ? out_img_stride : (long long)N * HW;
  a.out_plane_stride = out_plane_stride;
  a.hl_rows = hl_rows;
  a.out_f32 = out_f32;
  a.f32_img_stride = f32_img_stride > 0 ? f32_img_stride : (long long)(N - hl_rows) * HW;
  a.mask_bits = static_cast<const uint32_t*>(mask_bits);
  a.mask_img_stride = mask_img_stride > 0 ? mask_img_stride : (long long)N * (HW / 32);
  a.mask_rows = mask_rows > 0 ? mask_rows : N;
  a.mask_out = static_cast<uint32_t*>(mask_out);
  a.mask_out_img_stride = mask_out_img_stride > 0 ? mask_out_img_stride : (long long)N * (HW / 32);
  a.mask_out_rows = mask_out_rows > 0 ? mask_out_rows : N;
  if (planes == 2 && hl_rows > 0 && out_plane_stride <= 0) return fail(GNRF_ERR_ARG, "lin_hl: out_plane_stride required with 2 planes");
  CUtensorMap out_map = map;   // unused when no row is written as planes
  if (hl_rows > 0) {
    const int rows_pl = hl_rows < N ? hl_rows : N;
    const cuuint64_t gdim[4] = {(cuuint64_t)HW, (cuuint64_t)rows_pl, (cuuint64_t)n_img, (cuuint64_t)planes};
    const cuuint64_t gstr[3] = {(cuuint64_t)HW * 2, (cuuint64_t)a.out_img_stride * 2,
                                (cuuint64_t)(planes == 2 ? out_plane_stride : a.out_img_stride * n_img) * 2};
    const cuuint32_t box[4] = {32, 32, 1, (cuuint32_t)planes}, estr[4] = {1, 1, 1, 1};
    CUresult cr = enc(&out_map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, out, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return fail(GNRF_ERR_CUDA, "lin_hl: cuTensorMapEncodeTiled(out) failed (%d)", (int)cr);
  }
  int rc = planes == 2 ? tc::lin_hl_launch_t<2>(a, map, out_map, n_sm, as_stream(stream))
                       : tc::lin_hl_launch_t<1>(a, map, out_map, n_sm, as_stream(stream));
  if (rc != GNRF_OK) return rc;
  GNRF_LAUNCH_CHECK();
  count_launches(1);
  return GNRF_OK;
}

extern "C" size_t gnrf_wgrad_hl_workspace_bytes(int N, int K, int n_img, int HW) {
  if (N <= 0 || K <= 0 || n_img <= 0 || HW <= 0) return 0;
  return tc::wgrad_hl_plan(N, K, n_img, HW).partial_bytes;
}

extern "C" int gnrf_wgrad_hl(const void* dY, long long dy_img_stride, long long dy_plane_stride, const void* X, long long x_img_stride,
                             long long x_plane_stride, int planes, int N, int K, int n_img, int HW, float* dW, float* db, int db_sum,
                             void* workspace, size_t workspace_bytes, gnrf_stream_t stream) {
  GNRF_CHECK_ARG(dY && X && dW && workspace && N > 0 && K > 0 && n_img > 0 && HW > 0 && (planes == 1 || planes == 2));
  GNRF_CHECK_ARG(HW % tc::kWhKb == 0 && N >= 128);
  GNRF_CHECK_ARG((reinterpret_cast<uintptr_t>(dY) & 15) == 0 && (reinterpret_cast<uintptr_t>(X) & 15) == 0 && HW % 8 == 0);
  if (dy_img_stride <= 0) dy_img_stride = (long long)N * HW;
  if (x_img_stride <= 0) x_img_stride = (long long)K * HW;
  GNRF_CHECK_ARG(dy_img_stride % 8 == 0 && x_img_stride % 8 == 0 && dy_plane_stride % 8 == 0 && x_plane_stride % 8 == 0);
  if (planes == 2 && (dy_plane_stride <= 0 || x_plane_stride <= 0)) return fail(GNRF_ERR_ARG, "wgrad_hl: plane strides required with 2 planes");
  int n_sm = 0;
  {
    int rc = device_once(kOnceLinHlSm, &n_sm, []() -> int { return GNRF_OK; });
    if (rc != GNRF_OK) return rc;
  }
  const tc::WgHlPlan pl = tc::wgrad_hl_plan(N, K, n_img, HW);
  if (workspace_bytes < pl.partial_bytes) return fail(GNRF_ERR_ARG, "wgrad_hl: workspace %zu < required %zu bytes", workspace_bytes, pl.partial_bytes);
  tc::TensorMapEncodeFn enc = tc::hl_tensor_map_encoder();
  if (enc == nullptr) return fail(GNRF_ERR_CUDA, "wgrad_hl: cuTensorMapEncodeTiled is not available from the driver");
  CUtensorMap dy_map, x_map;
  {
    const cuuint64_t gdim[4] = {(cuuint64_t)HW, (cuuint64_t)N, (cuuint64_t)n_img, (cuuint64_t)planes};
    const cuuint64_t gstr[3] = {(cuuint64_t)HW * 2, (cuuint64_t)dy_img_stride * 2,
                                (cuuint64_t)(planes == 2 ? dy_plane_stride : dy_img_stride * n_img) * 2};
    const cuuint32_t box[4] = {(cuuint32_t)tc::kWhKb, 128, 1, (cuuint32_t)planes}, estr[4] = {1, 1, 1, 1};
    CUresult cr = enc(&dy_map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(dY), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return fail(GNRF_ERR_CUDA, "wgrad_hl: cuTensorMapEncodeTiled(dY) failed (%d)", (int)cr);
  }
  {
    const cuuint64_t gdim[4] = {(cuuint64_t)HW, (cuuint64_t)K, (cuuint64_t)n_img, (cuuint64_t)planes};
    const cuuint64_t gstr[3] = {(cuuint64_t)HW * 2, (cuuint64_t)x_img_stride * 2,
                                (cuuint64_t)(planes == 2 ? x_plane_stride : x_img_stride * n_img) * 2};
    const cuuint32_t box[4] = {(cuuint32_t)tc::kWhKb, (cuuint32_t)pl.box_rows, 1, 1}, estr[4] = {1, 1, 1, 1};
    CUresult cr = enc(&x_map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(X), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return fail(GNRF_ERR_CUDA, "wgrad_hl: cuTensorMapEncodeTiled(X) failed (%d)", (int)cr);
  }
  tc::WgHlArgs a;
  a.N_dy = N; a.K_x = K; a.n_img = n_img;
  a.n_mt = pl.n_mt; a.n_ch = pl.n_ch; a.chunk_rows = pl.chunk_rows; a.chunk_n = pl.chunk_n; a.n_box = pl.n_box; a.box_rows = pl.box_rows;
  a.n_split = pl.n_split; a.n_kb_total = pl.n_kb_total; a.n_items = pl.n_items; a.nb0 = pl.nb0; a.nb1 = pl.nb1;
  a.stage_tx_bytes = planes * (tc::kWhAPlane + pl.n_box * pl.box_rows * 64);
  a.partial = static_cast<float*>(workspace);
  int rc = planes == 2 ? tc::wgrad_hl_launch_t<2>(a, dy_map, x_map, n_sm, as_stream(stream))
                       : tc::wgrad_hl_launch_t<1>(a, dy_map, x_map, n_sm, as_stream(stream));
  if (rc != GNRF_OK) return rc;
  const long long total = (long long)N * (K + 1);
  tc::wgrad_hl_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(
      a.partial, N, K, n_img, pl.n_mt, pl.n_ch, pl.n_split, pl.chunk_rows, pl.chunk_n, dW, db, db_sum);
  GNRF_LAUNCH_CHECK();
  count_launches(2);
  return GNRF_OK;
}
