// Fused radiance-MLP kernel for sm_100a (tcgen05 / TMEM / TMA bulk copies), both branches in one launch.
//
//   tile = 128 consecutive sample points of one face (= 128/N_s whole rays)  ->  UMMA M = 128 (TMEM lane == point)
//   per tile:  ray points -> positional encoding -> L0..L8 dense stack -> ReLU density -> per-ray alpha composite
//
// Precision: every product is evaluated as bf16x3 -- x = x_hi + x_lo, w = w_hi + w_lo (bf16 each) and
//   x*w ~= x_hi*w_hi + x_lo*w_hi + x_hi*w_lo   (3 UMMAs, fp32 accumulation in TMEM; the dropped lo*lo term is 2^-18 relative)
// which keeps ~16 significand bits per operand: measured <= 2e-5 relative on the composited features vs the fp32 oracle
// (single-pass bf16 is 5e-3..9e-3 and TF32 4e-4..1.1e-3, SURVEY §7 hard part 1 -- outside / on the edge of the 1e-3 bar).
//
// Exact algebraic rewrites of the reference graph (models/mlp_nerf.py:95-119, models/gaze_nerf.py:248-262):
//   (i)   the 181 code channels of the 244-wide input and the 127 appearance channels are constant per face: their
//         weight columns are folded into per-face bias vectors (gnrf_mlp_tc_fold);  per-point K becomes 64 / 384 / 384+64.
//   (ii)  density_module shares its input with RGB_layer_0 -> appended as one extra output row of the last stage.
//   (iii) RGB_layer_0 -> RGB_layer_1 has no activation in between -> pre-multiplied at pack time (fp64) into one 192x384 layer.
//   (iv)  RGB_layer_2 is linear and un-activated: sum_s w_s (W h_s + b) = W (sum_s w_s h_s) + b sum_s w_s, so the 192-d hidden
//         is composited per ray and RGB_layer_2 is applied once per ray by rgb_head_kernel (fp32 FMA).
//
// On-chip layout per CTA (1 CTA / SM, 320 threads = 8 epilogue warps + 1 TMA warp + 1 MMA warp):
//   smem  A_hi       : 6 K-blocks x [128 rows x 64 bf16], SWIZZLE_128B K-major (UMMA canonical)          96 KB
//         A_lo[4..5] : K-blocks 4,5 of the low halves, same layout                                        32 KB
//         W ring     : 4 x 24 KB stages; a stage = one K-slice of one accumulator block, W_hi then W_lo (no-swizzle
//                      K-major core-matrix layout), streamed from the pre-arranged packed image in L2 by cp.async.bulk
//                      (TMA bulk copy) + mbarrier complete_tx                                              96 KB
//   TMEM  columns 0..383   : fp32 accumulators (block 0 = cols 0..191, block 1 = cols 192..383)
//         columns 384..511 : A_lo K-blocks 0..3 as packed bf16 pairs (32 columns per K-block) -> TS-mode UMMA reads the A
//                      operand straight from tensor memory; this is what frees 64 KB of smem for a ring deep enough to
//                      cover the L2 -> smem latency (profiles/: the 32 KB ring version was latency/issue bound at 39 %).
// Layer pipeline: the epilogue warps drain an N=384 accumulator K-block by K-block (64 columns: +bias, ReLU, hi/lo split,
// swizzled st.shared / tcgen05.st) and release each K-block to the MMA warp through its own mbarrier.  Both accumulator blocks
// are N=192 (96 tensor cycles per UMMA >= the single-thread issue cost, so both run at the tensor rate); block 0's columns
// (K-blocks 0..2 of the layer's output) are drained EARLY, while block 1 of the same layer is still running, so the next
// layer's block 0 starts the moment the layer completes and the late drain of K-blocks 3..5 stays ahead of it.
// The positional encoding needed again by layer 5 (skip connection) is recomputed by the epilogue warps while they would
// otherwise wait for layer 5's first pass.
#include <cstdlib>
#include <type_traits>

#include "common.cuh"
#include "sm100_ptx.cuh"
#include "tc_common.cuh"

namespace gnrf {
namespace tc {

using namespace ptx;

constexpr int kHidden = 384;
constexpr int kFeat = 258;
constexpr int kH2 = 192;                 // RGB_layer_1 width
constexpr int kTile = 128;               // points per tile == UMMA M
constexpr int kKB = 64;                  // K-block (bf16 elements) == 128-byte swizzle span
constexpr int kNumKB = kHidden / kKB;    // 6
constexpr int kL8N = 208;                // 192 + 1 density + 15 zero rows
constexpr int kSlotBytes = 24576;        // W ring slot (W_hi slices + W_lo slices of one stage)
constexpr int kStages = 4;
constexpr int kALoTmemKB = 4;            // A_lo K-blocks 0..3 live in TMEM, 4..5 in smem
constexpr int kALoCol = 384;             // first TMEM column of A_lo
constexpr int kABlockBytes = kTile * 128;            // 16384: one K-block of A (hi or lo)
constexpr int kABytes = kNumKB * kABlockBytes;       // 98304
constexpr int kEpiThreads = 256;         // 8 epilogue warps: warp w drains TMEM lanes 32*(w%4).. and column half (w/4) of each K-block
constexpr int kThreads = 320;            // + 1 TMA producer warp (8) + 1 MMA issue warp (9)
constexpr int kMaxGrid = 160;            // persistent grid upper bound (>= SM count of a B200: 148)
constexpr int kScratchLd = 193;          // composite scratch row stride (floats): conflict-free for both phases

// ---- weight stage schedule -------------------------------------------------------------------------------------------
// Every layer's N = 384 is split into two accumulator blocks of N = 192: blk 0 = outputs [0,192) (TMEM cols 0..191) and
// blk 1 = outputs [192,384) (cols 192..383); the last stage (192 RGB rows + density + pad) is ONE N = 208 block.  blk 0 goes first:
// it only needs its own 192 columns (K-blocks 0..2) drained, which the epilogue does early (see the kernel).  History (profiles/,
// tests/ubench): with 128 + 256 blocks the N=128 UMMAs (64 tensor cycles) were bound by the ~88-cycle single-thread issue cost
// (blk 0 took 6.7 K cycles instead of 4.6 K per layer); at N = 192 every UMMA covers 96 tensor cycles.
// A ring stage is a K-slice of one block: [nrows x 16] bf16 per K16 step in the no-swizzle K-major core-matrix layout
// (8 rows x 16 B contiguous, K-halves 128 B apart, 8-row groups 256 B apart), the W_hi slice followed by the W_lo slice;
// a stage is 2 K16 steps of an N=192 block = 24 KB = 6 UMMAs = 576 tensor cycles (last stage: 1 K16 step of N=208, 13 KB).
// Stream == consumption order:
//   for layer: for phase (layer 5 only: hidden columns, then PE columns): for blk: for K-slice: one stage
struct StageInfo {
  int layer, phase2, blk, ks, nk16, n0, nrows;
  uint32_t bytes;  // hi + lo
};
template <class F>
__host__ __device__ inline void for_each_stage(F&& f) {
  for (int layer = 0; layer < 9; ++layer) {
    const int n_phase = (layer == 5) ? 2 : 1;
    for (int ph = 0; ph < n_phase; ++ph) {
      const int k16n = (layer == 0 || ph == 1) ? 4 : 24;
      const int n_blk = layer == 8 ? 1 : 2;
      for (int blk = 0; blk < n_blk; ++blk) {
        const int n0 = blk ? 192 : 0;
        const int nrows = layer == 8 ? kL8N : 192;
        const int step = layer == 8 ? 1 : 2;   // K16 steps per stage
        for (int ks = 0; ks < k16n; ks += step) {
          StageInfo st{layer, ph, blk, ks, step, n0, nrows, (uint32_t)(2 * nrows * 32 * step)};
          f(st);
        }
      }
    }
  }
}
constexpr int stage_count() {
  int n = 0;
  for (int layer = 0; layer < 9; ++layer)
    for (int ph = 0; ph < ((layer == 5) ? 2 : 1); ++ph) {
      const int k16n = (layer == 0 || ph == 1) ? 4 : 24;
      n += layer == 8 ? k16n : 2 * (k16n / 2);
    }
  return n;
}
constexpr size_t stream_bytes() {
  size_t n = 0;
  for (int layer = 0; layer < 9; ++layer)
    for (int ph = 0; ph < ((layer == 5) ? 2 : 1); ++ph) {
      const int k16n = (layer == 0 || ph == 1) ? 4 : 24;
      n += (size_t)k16n * 2 * 32 * (layer == 8 ? kL8N : 384);
    }
  return n;
}
constexpr int kNumStagesPerTile = stage_count();   // 200
constexpr size_t kStreamBytes = stream_bytes();
// fp32 auxiliary block (float offsets from aux base)
constexpr int kBiasFloats = 8 * kHidden + kL8N;                       // 3280 per face
constexpr size_t kAuxBaseBias = 0;                                    // [3280] face-independent part of every bias
constexpr size_t kAuxW0c = kAuxBaseBias + kBiasFloats;                // [384][181] code columns of FeaExt_module_0
constexpr size_t kAuxW5c = kAuxW0c + (size_t)kHidden * GNRF_SHAPE_EXT_DIMS;
constexpr size_t kAuxW1c = kAuxW5c + (size_t)kHidden * GNRF_SHAPE_EXT_DIMS;   // [192][127] appearance columns of RGB_layer_1
constexpr size_t kAuxW2t = kAuxW1c + (size_t)kH2 * GNRF_APPEA_DIMS;           // [192][258] RGB_layer_2 transposed
constexpr size_t kAuxB2 = kAuxW2t + (size_t)kH2 * kFeat;                      // [258]
constexpr size_t kAuxWf = kAuxB2 + 264;                                       // [208][384] fused last-stage matrix (fp32)
constexpr int kVdDims = 27;                                                   // view-direction encoding: 3 + 6 * 4 (include_vd=True)
constexpr size_t kAuxW1v = kAuxWf + (size_t)kL8N * kHidden;                   // [192][27] view-direction columns of RGB_layer_1 (or zeros)
constexpr size_t kAuxFloats = kAuxW1v + (size_t)kH2 * 32;
constexpr size_t kPackedBytes = kStreamBytes + kAuxFloats * sizeof(float);

__host__ __device__ inline int bias_offset(int layer) { return layer * kHidden; }

// ---- shared memory map ---------------------------------------------------------------------------------------------
constexpr int kSmemAHi = 0;
constexpr int kSmemALo45 = kABytes;                                         // A_lo K-blocks 4,5
constexpr int kSmemRing = kABytes + (kNumKB - kALoTmemKB) * kABlockBytes;   // 131072
constexpr int kSmemBars = kSmemRing + kStages * kSlotBytes;                 // 229376
constexpr int kBarWFull = 0, kBarWEmpty = kStages, kBarAReady = 2 * kStages, kBarAccFull = kBarAReady + kNumKB,
              kBarAFree = kBarAccFull + 1, kBarAEarly = kBarAFree + 1, kNumBars = kBarAEarly + 1;
constexpr int kSmemMisc = kSmemBars + kNumBars * 8;          // tmem ptr, scan scratch
constexpr int kSmemBytes = kSmemMisc + 64 + 1024;            // + alignment slack

struct BranchArgs {
  const unsigned char* stream;  // packed bf16 stage stream
  const float* vd_bias;         // [B][N_r][192] per-RAY bias of the last stage (include_vd=True: view-direction columns) or null
  const float* bias;            // [B][kBiasFloats]
  float* hc;                    // [B][N_r][192]  composited hidden
  float* wsum;                  // [B][N_r]
  float* weights;               // [B][N_r][N_s] or null
};

struct FwdArgs {
  BranchArgs br[2];
  const float4* ray_dl;
  const float* tvecs;
  const float* z_edges;
  int n_branch, B, N_r, N_s, tiles_per_face, n_items;
  uint32_t* pe_stash; // [grid][65][128] u32 per-CTA stash of the tile's packed positional encoding (+ delta), L2 resident
  float* dbg;        // optional [10][128][384] dump of tile 0 activations
  long long* prof;   // optional timeline of CTA 0: [item < 4][layer 0..9][16] clock64 stamps / stall sums
};

// =====================================================================================================================
//  main kernel
// =====================================================================================================================
// One thread writes HALF a row (32 K values = 64 bytes hi + 64 bytes lo) of K-block kb: hi -> smem chunks 4*half..4*half+3
// (16-byte chunk index XOR (row & 7)); lo -> TMEM columns (kb < 4; 16 packed columns) or smem (kb 4,5).
__device__ __forceinline__ void st_shared_row64(uint32_t addr_row, uint32_t sw, int half, const uint32_t (&w)[16]) {
#pragma unroll
  for (int j = 0; j < 4; ++j)
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr_row + ((((uint32_t)(4 * half + j)) ^ sw) << 4)), "r"(w[4 * j]),
                 "r"(w[4 * j + 1]), "r"(w[4 * j + 2]), "r"(w[4 * j + 3])
                 : "memory");
}
__device__ __forceinline__ void store_a_half(uint32_t smem_base, uint32_t t_lane, int kb, int row, int half, const uint32_t (&hi)[16],
                                             const uint32_t (&lo)[16]) {
  const uint32_t row_off = a_row_offset(row);
  const uint32_t sw = (uint32_t)(row & 7);
  st_shared_row64(smem_base + kSmemAHi + (uint32_t)kb * kABlockBytes + row_off, sw, half, hi);
  if (kb < kALoTmemKB) {
    tmem_st16(t_lane + kALoCol + kb * 32 + half * 16, lo);   // completion is awaited once per release (tcgen05.wait::st)
  } else {
    st_shared_row64(smem_base + kSmemALo45 + (uint32_t)(kb - kALoTmemKB) * kABlockBytes + row_off, sw, half, lo);
  }
}

// Positional encoding of one sample point, K order = [x,y,z, f0: sin xyz, cos xyz, f1: ...] + zero pad (63 -> 64).
// utils/model_utils.py:263-280; pts = o + ((d * l) * z) with individually rounded ops (:315).
__device__ __forceinline__ void compute_pe(const float* __restrict__ tvec, const float4 dl, float z, float (&pe)[64]) {
  const float dc[3] = {dl.x, dl.y, dl.z};
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float x = __fadd_rn(tvec[c], __fmul_rn(__fmul_rn(dc[c], dl.w), z));
    pe[c] = x;
    float f = 1.0f;
#pragma unroll
    for (int q = 0; q < 10; ++q) {
      float sv, cv;
      sincosf(__fmul_rn(x, f), &sv, &cv);  // freq = 2^q exactly
      pe[3 + 6 * q + c] = sv;
      pe[6 + 6 * q + c] = cv;
      f *= 2.0f;
    }
  }
  pe[63] = 0.0f;
}

template <int CSIZE, bool PROF>
__global__ void __launch_bounds__(kThreads, 1) mlp_tc_kernel(const FwdArgs args) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  unsigned char* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bars = smem_base + kSmemBars;
  auto bar = [&](int i) { return bars + (uint32_t)i * 8u; };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem_gen + kSmemMisc);
  float* warp_prod = reinterpret_cast<float*>(smem_gen + kSmemMisc + 16);  // [4]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // Thread-block cluster (1, 2 or 4 CTAs): the CTAs of a cluster work on tiles of the SAME branch in lockstep; each loads 1/csize
  // of every weight stage from L2 and multicasts it to all of them (the weight feed was L2-bandwidth bound, profiles/).
  constexpr int csize = CSIZE;
  const int crank = (CSIZE > 1) ? (int)cluster_ctarank() : 0;
  const uint16_t cmask = (uint16_t)((1u << csize) - 1u);
  const int n_citems = args.n_items / csize;           // cluster work items
  const int cluster_id = (int)blockIdx.x / csize, n_clusters = (int)gridDim.x / csize;
  // cluster item -> (branch, face, tile) of THIS CTA
  auto decode_item = [&](int ci, int& branch, int& b, int& tile) {
    branch = ci % args.n_branch;
    const int t = (ci / args.n_branch) * csize + crank;
    b = t / args.tiles_per_face;
    tile = t - b * args.tiles_per_face;
  };

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(bar(kBarWFull + i), 1);
      mbar_init(bar(kBarWEmpty + i), (uint32_t)csize);   // released by the MMA warp of every CTA of the cluster
    }
    for (int i = 0; i < kNumKB; ++i) mbar_init(bar(kBarAReady + i), 8);  // one arrive per epilogue warp
    mbar_init(bar(kBarAccFull), 1);
    mbar_init(bar(kBarAFree), 1);
    mbar_init(bar(kBarAEarly), 1);
    fence_mbar_init();
  }
  if (warp == 9) tmem_alloc_512(smem_u32(const_cast<uint32_t*>(tmem_ptr_smem)));
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();   // every CTA's barriers are initialised before any peer signals them
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 8) {
    // =============================================== TMA producer ===============================================
    if (elect_one()) {
      uint32_t slot = 0, phase = 0;
      for (int ci = cluster_id; ci < n_citems; ci += n_clusters) {
        const int branch = ci % args.n_branch;
        const unsigned char* src = args.br[branch].stream;
        for_each_stage([&](const StageInfo& st) {  // strictly in consumption order
          mbar_wait_spin(bar(kBarWEmpty + slot), phase ^ 1);           // slot free in EVERY CTA of the cluster
          mbar_arrive_expect_tx(bar(kBarWFull + slot), st.bytes);       // the whole stage: own slice + the peers' multicasts
          const uint32_t part = st.bytes / (uint32_t)csize;
          const uint32_t dst = smem_base + kSmemRing + slot * kSlotBytes + (uint32_t)crank * part;
          if (csize == 1) bulk_g2s(dst, src, part, bar(kBarWFull + slot));
          else bulk_g2s_multicast(dst, src + (size_t)crank * part, part, bar(kBarWFull + slot), cmask);
          src += st.bytes;
          if (++slot == kStages) { slot = 0; phase ^= 1; }
        });
      }
    }
    __syncwarp();
  } else if (warp == 9) {
    // =============================================== MMA issuer =================================================
    // ONE elected lane runs the whole issue loop (measured, tests/ubench/ub_stage.cu: a per-stage elect + __syncwarp in a
    // converged-warp loop costs ~136 cycles per stage, which bursts of 64-cycle N=128 UMMAs cannot hide; inside a single
    // elect.sync region the compiler keeps descriptors in uniform registers and the loop runs at the tensor-pipe rate).
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t sbase_u = __shfl_sync(0xffffffffu, smem_base, 0);
    if (elect_one()) {
      uint32_t wp = 0;       // ring-wrap parity at the start of the current pass
      uint32_t a_phase = 0;  // bit kb = parity of the next completion of a_ready[kb]
      bool prof_on = false;
      long long w_stall = 0, a_stall = 0;
      int prof_item = 0, prof_layer = 0;
      auto stamp = [&](int layer, int k) {
        if (PROF) {
          prof_layer = layer;
          if (prof_on) {
            long long* p = args.prof + ((size_t)prof_item * 10 + layer) * 16;
            p[k] = clock64(); p[k + 1] = w_stall; p[k + 2] = a_stall;
          }
        }
      };
      constexpr uint32_t kDescHiSw128 = (uint32_t)((1024 >> 4) | (1u << 14) | (2u << 29));   // SBO, version, SWIZZLE_128B
      constexpr uint32_t kDescHiNoSw = (uint32_t)((256 >> 4) | (1u << 14));                  // SBO, version, no swizzle
      constexpr uint32_t kDescLoLboSw = 1u << 16, kDescLoLboNo = (128u >> 4) << 16;
      const uint32_t a_hi_lo0 = (((sbase_u + kSmemAHi) >> 4) & 0x3FFFu) | kDescLoLboSw;
      const uint32_t a_lo45_lo0 = (((sbase_u + kSmemALo45) >> 4) & 0x3FFFu) | kDescLoLboSw;
      const uint32_t ring_lo0 = (((sbase_u + kSmemRing) >> 4) & 0x3FFFu) | kDescLoLboNo;
      auto mk = [](uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; };

      // One ring stage = NK16 K16-steps of one accumulator block: per step  A_hi*W_hi, A_hi*W_lo, A_lo*W_hi.
      // `j` (stage index inside the pass) and `a_k16` are compile-time constants after unrolling: every pass has a multiple of
      // kStages stages, so slot = j % kStages and all descriptors fold to base + constant (the issue thread is the bottleneck
      // for 64-cycle N=128 UMMAs; tests/ubench/ub_pattern.cu shows the pattern itself runs at the ideal rate).
      auto do_stage = [&](auto nrows_c, auto nk16_c, int j, uint32_t d_col, int a_k16, bool fresh) {
        constexpr int NROWS = decltype(nrows_c)::value, NK16 = decltype(nk16_c)::value;
        constexpr uint32_t idesc = umma_idesc_bf16(128, NROWS);
        constexpr uint32_t kSliceU = (uint32_t)((NROWS * 32) >> 4);  // one K16 slice, in 16-byte units
        const int slot = j % kStages;
        const uint32_t ph = wp ^ (uint32_t)((j / kStages) & 1);
        long long tw0 = 0;
        if (PROF) tw0 = prof_on ? clock64() : 0;
        mbar_wait_spin(bar(kBarWFull + slot), ph);
        if (PROF) { if (prof_on) w_stall += clock64() - tw0; }
        tc_fence_after_sync();
        const uint32_t b_hi0 = ring_lo0 + (uint32_t)((slot * kSlotBytes) >> 4);
        const uint32_t b_lo0 = b_hi0 + NK16 * kSliceU;
        const uint32_t d = tmem_u + d_col;
#pragma unroll
        for (int i = 0; i < NK16; ++i) {
          const int k16 = a_k16 + i;
          const int kb = k16 >> 2;
          // A_hi: SW128 K-block kb, +32 B per K16 step inside the 128-byte swizzle span
          const uint64_t a_hi = mk(a_hi_lo0 + (uint32_t)((kb * kABlockBytes + (k16 & 3) * 32) >> 4), kDescHiSw128);
          const uint64_t b_hi = mk(b_hi0 + i * kSliceU, kDescHiNoSw);
          const uint64_t b_lo = mk(b_lo0 + i * kSliceU, kDescHiNoSw);
          // A_hi is used by two consecutive UMMAs: read it from smem once (A collector fill / lastuse)
          umma_ss_a_fill(d, a_hi, b_hi, idesc, (fresh && i == 0) ? 0u : 1u);
          umma_ss_a_lastuse(d, a_hi, b_lo, idesc, 1u);
          if (kb < kALoTmemKB)
            umma_ts(d, tmem_u + (uint32_t)(kALoCol + kb * 32 + (k16 & 3) * 8), b_hi, idesc, 1u);
          else
            umma_ss(d, mk(a_lo45_lo0 + (uint32_t)(((kb - kALoTmemKB) * kABlockBytes + (k16 & 3) * 32) >> 4), kDescHiSw128), b_hi,
                    idesc, 1u);
        }
        if (CSIZE == 1) umma_commit(bar(kBarWEmpty + slot));
        else umma_commit_multicast(bar(kBarWEmpty + slot), cmask);
      };
      auto wait_a = [&](int kb) {
        long long ta0 = 0;
        if (PROF) ta0 = prof_on ? clock64() : 0;
        mbar_wait_spin(bar(kBarAReady + kb), (a_phase >> kb) & 1u);
        if (PROF) { if (prof_on) a_stall += clock64() - ta0; }
        a_phase ^= (1u << kb);
        tc_fence_after_sync();
      };
      // all K-slices of both accumulator blocks (N = 192 each: 96 tensor cycles per UMMA, above the ~88-cycle single-thread issue
      // cost, so both blocks run at the tensor rate), fully unrolled; a stage = 2 K16 steps = 6 UMMAs.
      // `early`: accumulator columns 0..191 (blk 0) are final once blk 0 has been issued, and A K-blocks 0..2 are dead once blk 1 has
      // consumed its first six stages -> commit kBarAEarly there, so the epilogue drains K-blocks 0..2 of this layer's output (the
      // first operands AND the accumulator columns of the next layer's blk 0) while blk 1 is still running.
      // `afree`: (layer 5) A K-block 0 is dead once blk 1 has consumed its first two stages; kBarAFree is committed there, so the
      // epilogue re-stages the positional encoding into K-block 0 while the rest of the hidden pass is still running.
      auto run_pass = [&](auto k16n_c, bool pe_pass, bool fresh_start, bool early, bool afree = false) {
        constexpr int K16N = decltype(k16n_c)::value;
        constexpr int NBH = K16N / 2;   // stages per block
        static_assert((2 * NBH) % kStages == 0, "a pass must use a whole number of ring wraps");
        using N192 = std::integral_constant<int, 192>;
        using S2 = std::integral_constant<int, 2>;
#pragma unroll
        for (int i = 0; i < NBH; ++i) {
          // blk 0's first UMMA overwrites columns 0..191 = K-blocks 0..2 of the previous accumulator: all three must have been
          // drained before it is issued; K-blocks 3..5 are awaited when blk 0 first reads them (two stages per K-block).
          if (!pe_pass) {
            if (i == 0) { wait_a(0); wait_a(1); wait_a(2); }
            else if (i >= 6 && (i & 1) == 0) wait_a(i >> 1);
          }
          do_stage(N192{}, S2{}, i, 0u, 2 * i, fresh_start && i == 0);
          if (PROF) { if (prof_on && K16N == 24 && i >= 2 && (i & 1) == 1) args.prof[((size_t)prof_item * 10 + prof_layer) * 16 + 10 + (i >> 1)] = clock64(); }
        }
        if (PROF) { if (prof_on && K16N == 24) args.prof[((size_t)prof_item * 10 + prof_layer) * 16 + 10] = clock64(); }  // blk 0 issued
#pragma unroll
        for (int i = 0; i < NBH; ++i) {
          do_stage(N192{}, S2{}, NBH + i, 192u, 2 * i, fresh_start && i == 0);
          if (K16N == 24 && i == 5 && early) umma_commit(bar(kBarAEarly));
          if (K16N == 24 && i == 1 && afree) umma_commit(bar(kBarAFree));
        }
        if (((2 * NBH) / kStages) & 1) wp ^= 1u;
      };
      // last stage: one N = 208 block (192 RGB rows + density + pad), one K16 step per ring stage
      auto run_last = [&]() {
        using N208 = std::integral_constant<int, kL8N>;
        using S1 = std::integral_constant<int, 1>;
        static_assert(24 % kStages == 0, "a pass must use a whole number of ring wraps");
#pragma unroll
        for (int i = 0; i < 24; ++i) {
          if (i == 0) { wait_a(0); wait_a(1); wait_a(2); wait_a(3); }   // columns 0..207 drained, K-blocks 0..3 staged
          else if (i == 16) wait_a(4);
          else if (i == 20) wait_a(5);
          do_stage(N208{}, S1{}, i, 0u, i, i == 0);
        }
        if ((24 / kStages) & 1) wp ^= 1u;
      };
      using I4 = std::integral_constant<int, 4>;
      using I24 = std::integral_constant<int, 24>;

      for (int ci = cluster_id; ci < n_citems; ci += n_clusters) {
        if (PROF) {
          prof_item = ci / n_clusters;
          prof_on = (args.prof != nullptr) && blockIdx.x == 0 && prof_item < 4;
        }
        // ---- layer 0: K = 64 (PE in K-block 0)
        stamp(0, 0);
        wait_a(0);
        run_pass(I4{}, true, true, false);
        umma_commit(bar(kBarAEarly));
        umma_commit(bar(kBarAccFull));
        stamp(0, 3);
        // ---- layers 1..7
#pragma unroll 1
        for (int layer = 1; layer < 8; ++layer) {
          stamp(layer, 0);
          run_pass(I24{}, false, true, layer != 5, layer == 5);
          if (layer == 5) {
            // skip connection: the PE columns of FeaExt_module_5 (models/mlp_nerf.py:106-107); the epilogue re-staged the tile's
            // PE into K-block 0 as soon as the hidden pass had finished reading that K-block (kBarAFree, committed inside the pass).
            wait_a(0);
            run_pass(I4{}, true, false, false);
            umma_commit(bar(kBarAEarly));
          }
          umma_commit(bar(kBarAccFull));
          stamp(layer, 3);
        }
        // ---- layer 8: [RGB_0*RGB_1 (192) | density (1) | pad] = N 128 + 80
        stamp(8, 0);
        run_last();
        umma_commit(bar(kBarAccFull));
        stamp(8, 3);
      }
    }
    __syncwarp();
  } else {
    // =============================================== epilogue warps (0..7) =======================================
    // thread = (row, half): row = TMEM lane = sample point of the tile; half selects 32 of the 64 columns of every K-block
    // (the drain is latency bound, two warps per scheduler hide each other's tcgen05.ld / conversion latencies)
    const int half = warp >> 2;
    const int row = (warp & 3) * 32 + lane;
    const uint32_t t_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t acc_phase = 0, afree_phase = 0, a01_phase = 0;
    const int N_s = args.N_s;
    const int rays_per_tile = kTile / N_s;
    // per-ray composite scratch [128][193] fp32: aliases A K-blocks 1.. (K-block 0 stays free for the next tile's PE)
    float* scratch = reinterpret_cast<float*>(smem_gen + kABlockBytes);
    // per-CTA stash of the tile's positional encoding, L2 resident: [64 words + delta][2 halves][128 rows]; a thread only ever
    // reads back what it wrote itself (half 0 keeps the hi words, half 1 the lo words), so no cross-thread ordering is needed
    uint32_t* stash = args.pe_stash + (size_t)blockIdx.x * (65 * kTile) + row;

    auto release_kb = [&](int kb) {
      fence_proxy_async_smem();   // generic-proxy st.shared -> visible to the tensor core (async proxy)
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(kBarAReady + kb));
    };
    // one fence for two K-blocks (the fence is a MEMBAR.ALL.CTA: ~300 cycles with 8 warps storing)
    auto release_kb_one = [&](int kb) {
      tmem_wait_st();
      fence_proxy_async_smem();
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(kBarAReady + kb));
    };
    auto release_kb_pair = [&](int kb_even) {
      tmem_wait_st();
      fence_proxy_async_smem();
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) { mbar_arrive(bar(kBarAReady + kb_even)); mbar_arrive(bar(kBarAReady + kb_even + 1)); }
    };
    // sample point of (item, row) -> positional encoding (+ delta); both halves compute it (idle window), each keeps its words
    auto stash_pe = [&](int ci) {
      int branch_, b, tile;
      decode_item(ci, branch_, b, tile);
      const int ray = tile * rays_per_tile + row / N_s;
      const int s = row - (row / N_s) * N_s;
      const float4 dl = args.ray_dl[(size_t)b * args.N_r + ray];
      const float* ze = args.z_edges + ((size_t)b * args.N_r + ray) * (N_s + 1);
      const float z = ze[s];
      const float delta = __fmul_rn(__fsub_rn(ze[s + 1], z), dl.w);  // (z_{k+1} - z_k) * l  (utils/model_utils.py:309-310)
      float pe[64];
      compute_pe(args.tvecs + b * 3, dl, z, pe);
#pragma unroll
      for (int q = 0; q < 32; ++q) {
        uint32_t hi, lo;
        split2(pe[2 * q], pe[2 * q + 1], hi, lo);
        stash[(size_t)(half * 32 + q) * kTile] = half ? lo : hi;
      }
      if (half == 0) stash[(size_t)64 * kTile] = __float_as_uint(delta);
      if (args.dbg != nullptr && ci == 0 && blockIdx.x == 0 && half == 0)
        for (int j = 0; j < 64; ++j) args.dbg[(size_t)row * kHidden + j] = pe[j];
      return delta;
    };
    // stash -> A K-block 0 (half 0: hi words -> smem row; half 1: lo words -> TMEM), then release it to the MMA warp
    auto stage_pe = [&]() {
      uint32_t w[32];
#pragma unroll
      for (int q = 0; q < 32; ++q) w[q] = stash[(size_t)(half * 32 + q) * kTile];
      if (half == 0) {
        st_shared_row128(smem_base + kSmemAHi + a_row_offset(row), (uint32_t)(row & 7), w);
      } else {
        tmem_st32(t_lane + kALoCol, w);
        tmem_wait_st();
      }
      release_kb(0);
    };

    float delta = 0.0f, delta_next = 0.0f;
    if (cluster_id < n_citems) {
      delta = stash_pe(cluster_id);
      stage_pe();
    }
    for (int item = cluster_id; item < n_citems; item += n_clusters) {
      int branch, b, tile;
      decode_item(item, branch, b, tile);
      const BranchArgs& br = args.br[branch];
      const float* bias = br.bias + (size_t)b * kBiasFloats;
      const bool dump = (args.dbg != nullptr) && (item == 0) && blockIdx.x == 0;
      const int next_item = item + n_clusters;
      const int ray = tile * rays_per_tile + row / N_s;
      const int s = row - (row / N_s) * N_s;

      // ---- trunk: drain layer l accumulators into A as the input of layer l+1 ----------------------------------
      const bool eprof = PROF && (args.prof != nullptr) && blockIdx.x == 0 && (item / n_clusters) < 4 && threadIdx.x == 0;
      long long* ep = eprof ? args.prof + ((size_t)(item / n_clusters) * 10) * 16 : nullptr;
      if (eprof) ep[6] = clock64();  // layer-0 row: K-block 0 (PE) of this tile released
      for (int layer = 0; layer < 8; ++layer) {
        const float* bl = bias + bias_offset(layer) + half * 32;
        // drain one K-block (64 accumulator columns; this thread: 32 of them) into the next layer's A operand
        auto drain_kb = [&](int kb) {
          uint32_t r0[32];
          tmem_ld32(t_lane + kb * 64 + half * 32, r0);
          const float4* b4 = reinterpret_cast<const float4*>(bl + kb * 64);
          float4 bb[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) bb[j] = __ldg(b4 + j);
          tmem_wait_ld();
          float v[32];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            v[4 * j + 0] = fmaxf(__uint_as_float(r0[4 * j + 0]) + bb[j].x, 0.0f);
            v[4 * j + 1] = fmaxf(__uint_as_float(r0[4 * j + 1]) + bb[j].y, 0.0f);
            v[4 * j + 2] = fmaxf(__uint_as_float(r0[4 * j + 2]) + bb[j].z, 0.0f);
            v[4 * j + 3] = fmaxf(__uint_as_float(r0[4 * j + 3]) + bb[j].w, 0.0f);
          }
          {
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) split2(v[2 * q], v[2 * q + 1], hi[q], lo[q]);
            store_a_half(smem_base, t_lane, kb, row, half, hi, lo);
          }
          if (dump)
            for (int j = 0; j < 32; ++j) args.dbg[((size_t)(layer + 1) * kTile + row) * kHidden + kb * 64 + half * 32 + j] = v[j];
        };
        // K-blocks 0..2: accumulator columns 0..191 are final and A K-blocks 0..2 are no longer read (kBarAEarly) while the rest of
        // the layer (blk 1) may still be running; the next layer's blk 0 can then start the moment this layer completes.
        mbar_wait(bar(kBarAEarly), a01_phase);
        a01_phase ^= 1;
        tc_fence_after_sync();
        drain_kb(0);
        drain_kb(1);
        release_kb_pair(0);
        drain_kb(2);
        release_kb_one(2);
        if (eprof) ep[layer * 16 + 8] = clock64();  // K-blocks 0..2 released
        mbar_wait(bar(kBarAccFull), acc_phase);
        if (eprof) ep[layer * 16 + 7] = clock64();  // accumulator of `layer` complete
        acc_phase ^= 1;
        tc_fence_after_sync();
        drain_kb(3);
        release_kb_one(3);   // blk 0 of the next layer reads K-block 3 first
        if (eprof && layer > 0) ep[layer * 16 + 6] = clock64();
        drain_kb(4);
        drain_kb(5);
        release_kb_pair(4);
        if (eprof) ep[layer * 16 + 9] = clock64();  // all released
        if (layer == 4) {
          // layer 5 = [hidden | PE] (skip connection, models/mlp_nerf.py:106-107): once the MMA warp has consumed the hidden
          // K-blocks, re-stage the tile's PE (from the stash) into K-block 0.
          mbar_wait(bar(kBarAFree), afree_phase);
          afree_phase ^= 1;
          tc_fence_after_sync();
          stage_pe();
        }
        // idle window (the MMA warp is busy with layer 7): prepare the next tile's positional encoding
        if (layer == 6 && next_item < n_citems) delta_next = stash_pe(next_item);
      }

      // ---- last stage: density -> alpha -> transmittance scan -> weights; composite the 192-d hidden per ray ----
      mbar_wait(bar(kBarAccFull), acc_phase);
      if (eprof) ep[8 * 16 + 7] = clock64();
      acc_phase ^= 1;
      tc_fence_after_sync();
      const float* b8 = bias + bias_offset(8);
      const float* vdb = br.vd_bias ? br.vd_bias + ((size_t)b * args.N_r + ray) * kH2 : nullptr;   // include_vd: W1[:, 384:411] PE4(d_ray)
      float w_k;
      {
        uint32_t rs;
        tmem_ld1(t_lane + kH2, rs);
        tmem_wait_ld();
        const float sigma = fmaxf(__uint_as_float(rs) + __ldg(b8 + kH2), 0.0f);          // models/mlp_nerf.py:115
        const float alpha = __fsub_rn(1.0f, expf(-__fmul_rn(sigma, delta)));             // utils/model_utils.py:500
        const float x = __fadd_rn(__fsub_rn(1.0f, alpha), 1e-10f);                       // :507
        // exclusive product over the samples of this ray that precede s (segmented scan; cumprod of :508-510)
        const int seg = N_s < 32 ? N_s : 32;
        const int lane_in_seg = lane & (seg - 1);
        float incl = x;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
          float up = __shfl_up_sync(0xffffffffu, incl, off);
          if (off < seg && lane_in_seg >= off) incl *= up;
        }
        float T = __shfl_up_sync(0xffffffffu, incl, 1);
        if (lane_in_seg == 0) T = 1.0f;
        if (N_s > 32) {
          if (lane == 31 && half == 0) warp_prod[warp & 3] = incl;
          named_bar_sync(1, kEpiThreads);
          const int warps_per_ray = N_s >> 5;
          const int wq = warp & 3;
          const int w0 = (wq / warps_per_ray) * warps_per_ray;
          for (int w = w0; w < wq; ++w) T *= warp_prod[w];
        }
        w_k = __fmul_rn(alpha, T);                                                        // :512
        if (half == 0 && br.weights != nullptr) br.weights[((size_t)b * args.N_r + ray) * N_s + s] = w_k;
        if (dump && half == 0) args.dbg[((size_t)9 * kTile + row) * kHidden + kH2] = sigma;
      }
      // all A reads of this tile are complete (acc_full): move w * ReLU(hidden) out of TMEM into the smem scratch
#pragma unroll 1
      for (int g = 0; g < 3; ++g) {
        const int c0 = half * 96 + g * 32;
        uint32_t r[32];
        tmem_ld32(t_lane + c0, r);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float h = fmaxf(__uint_as_float(r[j]) + __ldg(b8 + c0 + j) + (vdb ? __ldg(vdb + c0 + j) : 0.0f), 0.0f);   // ReLU after RGB_layer_1 (:112)
          scratch[row * kScratchLd + c0 + j] = w_k * h;
          if (dump) args.dbg[((size_t)9 * kTile + row) * kHidden + c0 + j] = h;
        }
      }
      if (half == 0) scratch[row * kScratchLd + kH2] = w_k;
      // the accumulators are drained: hand the next tile's layer 0 to the MMA warp BEFORE the cross-row reduction
      if (next_item < n_citems) stage_pe();
      else tc_fence_before_sync();
      named_bar_sync(1, kEpiThreads);
      {
        const int total = rays_per_tile * kScratchLd;
        for (int idx = (int)threadIdx.x; idx < total; idx += kEpiThreads) {
          const int rl = idx / kScratchLd, c = idx - rl * kScratchLd;
          const float* col = scratch + (size_t)rl * N_s * kScratchLd + c;
          float acc = 0.0f;
          for (int k = 0; k < N_s; ++k) acc += col[(size_t)k * kScratchLd];
          const size_t gray = (size_t)b * args.N_r + tile * rays_per_tile + rl;
          if (c < kH2) br.hc[gray * kH2 + c] = acc; else br.wsum[gray] = acc;
        }
      }
      named_bar_sync(1, kEpiThreads);  // scratch reads done before the next tile's layer-0 drain overwrites A K-blocks 1..
      delta = delta_next;
      if (eprof) ep[8 * 16 + 9] = clock64();  // composite done
    }
  }

  // ---- teardown ----------------------------------------------------------------------------------------------
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();   // no CTA may exit while a peer can still multicast into its smem or signal its barriers
  if (warp == 9) {
    tc_fence_after_sync();
    tmem_dealloc_512(tmem_base);
  }
}

// =====================================================================================================================
//  RGB_layer_2 after compositing (fold iv):  feat[b][c][ray] = sum_k W2[c][k] hc[b][ray][k] + b2[c] wsum ; bg_alpha = 1 - wsum
// =====================================================================================================================
// smem-tiled SGEMM: CTA = 64 rays x 64 output channels, thread = 4 rays x 4 channels, K = 192 in slabs of 16 (double buffered through
// registers).  M = rays (hc is [ray][192], K contiguous), N = channels (W2^T is [192][258], N contiguous); the stores are float4 along the
// rays of one channel, i.e. 256 contiguous bytes per half warp of the NCHW-like feature map [B][258][N_r].
constexpr int kHeadTM = 64, kHeadTN = 64, kHeadBK = 16;
struct HeadArgs {
  const float* aux[2];
  const float* hc[2];
  const float* wsum[2];
  float* feat_ray[2];
  float* bg_alpha[2];
};
__global__ void __launch_bounds__(256) rgb_head_kernel(const HeadArgs ha, int N_r, int n_branch) {
  __shared__ __align__(16) float As[2][kHeadBK][kHeadTM + 4];   // [k][ray]
  __shared__ __align__(16) float Bs[2][kHeadBK][kHeadTN];       // [k][channel]
  const int br = blockIdx.z % n_branch, b = blockIdx.z / n_branch;
  const float* __restrict__ aux = ha.aux[br];
  const float* __restrict__ hc = ha.hc[br] + (size_t)b * N_r * kH2;
  const float* __restrict__ wsum = ha.wsum[br] + (size_t)b * N_r;
  const float* __restrict__ w2t = aux + kAuxW2t;
  const int r0 = blockIdx.x * kHeadTM, n0 = blockIdx.y * kHeadTN;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  // loaders: A: thread -> (ray = tid / 4, 4 consecutive k = (tid % 4) * 4) one float4; B: 4 scalars (k = tid / 16 [+0], channel = (tid % 16) * 4 + j)
  const int a_ray = tid >> 2, a_k = (tid & 3) * 4;
  const int b_k = tid >> 4, b_c = (tid & 15) * 4;
  float4 a_st;
  float b_st[4];
  auto fetch = [&](int k0) {
    a_st = (r0 + a_ray < N_r) ? __ldg(reinterpret_cast<const float4*>(hc + (size_t)(r0 + a_ray) * kH2 + k0 + a_k)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < 4; ++j) b_st[j] = (n0 + b_c + j < kFeat) ? __ldg(w2t + (size_t)(k0 + b_k) * kFeat + n0 + b_c + j) : 0.0f;
  };
  auto commit = [&](int buf) {
    As[buf][a_k + 0][a_ray] = a_st.x; As[buf][a_k + 1][a_ray] = a_st.y; As[buf][a_k + 2][a_ray] = a_st.z; As[buf][a_k + 3][a_ray] = a_st.w;
    *reinterpret_cast<float4*>(&Bs[buf][b_k][b_c]) = make_float4(b_st[0], b_st[1], b_st[2], b_st[3]);
  };
  float acc[4][4];   // [channel][ray]
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
  fetch(0);
  commit(0);
  __syncthreads();
  constexpr int n_k = kH2 / kHeadBK;   // 12
  for (int kc = 0; kc < n_k; ++kc) {
    const int buf = kc & 1;
    if (kc + 1 < n_k) fetch((kc + 1) * kHeadBK);
#pragma unroll
    for (int kk = 0; kk < kHeadBK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[buf][kk][tx * 4]);
      const float4 w = *reinterpret_cast<const float4*>(&Bs[buf][kk][ty * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(wv[i], av[j], acc[i][j]);
    }
    if (kc + 1 < n_k) commit(buf ^ 1);
    __syncthreads();
  }
  const int ray = r0 + tx * 4;
  if (ray >= N_r) return;
  float ws[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) ws[j] = (ray + j < N_r) ? __ldg(wsum + ray + j) : 0.0f;
  if (blockIdx.y == 0 && ty == 0) {   // bg_alpha = 1 - sum_k w_k   (utils/model_utils.py:531-532)
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (ray + j < N_r) ha.bg_alpha[br][(size_t)b * N_r + ray + j] = 1.0f - ws[j];
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = n0 + ty * 4 + i;
    if (c >= kFeat) continue;
    const float b2 = __ldg(aux + kAuxB2 + c);
    float* dst = ha.feat_ray[br] + ((size_t)b * kFeat + c) * N_r + ray;
    const float v[4] = {fmaf(b2, ws[0], acc[i][0]), fmaf(b2, ws[1], acc[i][1]), fmaf(b2, ws[2], acc[i][2]), fmaf(b2, ws[3], acc[i][3])};
    if (ray + 3 < N_r && (N_r & 3) == 0) {
      *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (ray + j < N_r) dst[j] = v[j];
    }
  }
}

// =====================================================================================================================
//  packing
// =====================================================================================================================
struct PackSrc {
  const float* w[12];
  const float* b[12];
  int n_vd;   // 0, or 27: RGB_layer_1's input is [hidden 384 | view-direction encoding 27 | appearance 127] (models/gaze_nerf.py:140-141)
};

// Wf[n][k] (n < 192) = sum_j W1[n][j] * W0rgb[j][k] (fp64 accumulate);  row 192 = density weights;  rows 193.. = 0.
// base_bias[8*384 + n] (n < 192) = b1[n] + sum_j W1[n][j] * b0rgb[j];  [.. + 192] = b_density.
__global__ void fuse_head_kernel(PackSrc src, float* __restrict__ aux) {
  const int n = blockIdx.x;  // 0..207
  float* wf = aux + kAuxWf + (size_t)n * kHidden;
  float* base_bias = aux + kAuxBaseBias;
  const int ld1 = kHidden + src.n_vd + GNRF_APPEA_DIMS;
  if (n < kH2) {
    for (int k = threadIdx.x; k < kHidden; k += blockDim.x) {
      double acc = 0.0;
      for (int j = 0; j < kHidden; ++j) acc += (double)src.w[10][(size_t)n * ld1 + j] * (double)src.w[9][(size_t)j * kHidden + k];
      wf[k] = (float)acc;
    }
    if (threadIdx.x == 0) {
      double acc = (double)src.b[10][n];
      for (int j = 0; j < kHidden; ++j) acc += (double)src.w[10][(size_t)n * ld1 + j] * (double)src.b[9][j];
      base_bias[8 * kHidden + n] = (float)acc;
    }
  } else if (n == kH2) {
    for (int k = threadIdx.x; k < kHidden; k += blockDim.x) wf[k] = src.w[8][k];
    if (threadIdx.x == 0) base_bias[8 * kHidden + n] = src.b[8][0];
  } else {
    for (int k = threadIdx.x; k < kHidden; k += blockDim.x) wf[k] = 0.0f;
    if (threadIdx.x == 0) base_bias[8 * kHidden + n] = 0.0f;
  }
}

// copies: trunk biases, code / appearance columns, RGB_layer_2 transposed
__global__ void pack_aux_kernel(PackSrc src, float* __restrict__ aux) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int nth = gridDim.x * blockDim.x;
  const int vp = GNRF_PE_DIMS + GNRF_SHAPE_EXT_DIMS;  // 244
  for (int i = tid; i < 8 * kHidden; i += nth) aux[kAuxBaseBias + i] = src.b[i / kHidden][i % kHidden];
  for (int i = tid; i < kHidden * GNRF_SHAPE_EXT_DIMS; i += nth) {
    int n = i / GNRF_SHAPE_EXT_DIMS, k = i % GNRF_SHAPE_EXT_DIMS;
    aux[kAuxW0c + i] = src.w[0][(size_t)n * vp + GNRF_PE_DIMS + k];
    aux[kAuxW5c + i] = src.w[5][(size_t)n * (vp + kHidden) + GNRF_PE_DIMS + k];
  }
  const int ld1 = kHidden + src.n_vd + GNRF_APPEA_DIMS;
  for (int i = tid; i < kH2 * GNRF_APPEA_DIMS; i += nth) {
    int n = i / GNRF_APPEA_DIMS, k = i % GNRF_APPEA_DIMS;
    aux[kAuxW1c + i] = src.w[10][(size_t)n * ld1 + kHidden + src.n_vd + k];
  }
  for (int i = tid; i < kH2 * 32; i += nth) {
    int n = i / 32, k = i % 32;
    aux[kAuxW1v + i] = (k < src.n_vd) ? src.w[10][(size_t)n * ld1 + kHidden + k] : 0.0f;
  }
  for (int i = tid; i < kH2 * kFeat; i += nth) {
    int k = i / kFeat, c = i % kFeat;
    aux[kAuxW2t + i] = src.w[11][(size_t)c * kH2 + k];
  }
  for (int i = tid; i < kFeat; i += nth) aux[kAuxB2 + i] = src.b[11][i];
}

// Stage record of one ring stage; every CTA of pack_stream_kernel derives its own from for_each_stage (no device-global table, no
// per-process state: the ABI stays re-entrant across streams and devices).
struct StageRec {
  uint32_t byte_off;
  uint16_t ks, nk16, n0, nrows;
  uint8_t layer, phase2, pad0, pad1;
};

__device__ __forceinline__ float pack_src_value(const PackSrc& src, const float* wf, int layer, int phase2, int n, int k) {
  const int vp = GNRF_PE_DIMS + GNRF_SHAPE_EXT_DIMS;
  if (layer == 0) return (k < GNRF_PE_DIMS) ? src.w[0][(size_t)n * vp + k] : 0.0f;
  if (layer == 5) {
    const size_t ld = vp + kHidden;
    if (phase2) return (k < GNRF_PE_DIMS) ? src.w[5][(size_t)n * ld + k] : 0.0f;
    return src.w[5][(size_t)n * ld + vp + k];
  }
  if (layer == 8) return wf[(size_t)n * kHidden + k];
  return src.w[layer][(size_t)n * kHidden + k];
}

// One CTA per stage, one thread per 16-byte chunk (8 bf16 along K of one row).
// Chunk order inside a K16 slice of `nrows` rows = the no-swizzle K-major core-matrix layout:
//   chunk = (row / 8) * 16 + k_half * 8 + (row % 8)        (8 rows x 16 B contiguous; K-halves 128 B apart; groups 256 B apart)
__global__ void pack_stream_kernel(PackSrc src, const float* __restrict__ wf, unsigned char* __restrict__ stream) {
  __shared__ StageRec s_st;
  if (threadIdx.x == 0) {
    int i = 0;
    uint32_t off = 0;
    for_each_stage([&](const StageInfo& si) {
      if (i == (int)blockIdx.x) {
        s_st.byte_off = off;
        s_st.ks = (uint16_t)si.ks; s_st.nk16 = (uint16_t)si.nk16; s_st.n0 = (uint16_t)si.n0; s_st.nrows = (uint16_t)si.nrows;
        s_st.layer = (uint8_t)si.layer; s_st.phase2 = (uint8_t)si.phase2; s_st.pad0 = 0; s_st.pad1 = 0;
      }
      ++i;
      off += si.bytes;
    });
  }
  __syncthreads();
  const StageRec st = s_st;
  const int chunks_per_slice = st.nrows * 2;
  const int chunks_per_half = chunks_per_slice * st.nk16;   // W_hi slices first, then the W_lo slices
  for (int c = threadIdx.x; c < 2 * chunks_per_half; c += blockDim.x) {
    const int half = c >= chunks_per_half;
    const int ch = c - half * chunks_per_half;
    const int slice = ch / chunks_per_slice;
    const int rem = ch - slice * chunks_per_slice;
    const int row = (rem >> 4) * 8 + (rem & 7);
    const int k_half = (rem >> 3) & 1;
    const int n = st.n0 + row;
    const int k0 = (st.ks + slice) * 16 + k_half * 8;
    uint32_t out[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float a = pack_src_value(src, wf, st.layer, st.phase2, n, k0 + 2 * q);
      float b = pack_src_value(src, wf, st.layer, st.phase2, n, k0 + 2 * q + 1);
      uint32_t hi, lo;
      split2(a, b, hi, lo);
      out[q] = half ? lo : hi;
    }
    *reinterpret_cast<uint4*>(stream + st.byte_off + (size_t)c * 16) = make_uint4(out[0], out[1], out[2], out[3]);
  }
}

// include_vd=True: per-RAY bias of the last stage  vd_bias[ray][n] = sum_j W1[n][384 + j] * PE4(d_ray)[j]  (n < 192), where PE4 is the
// reference's Embedder with 4 frequencies + the input on the normalised ray direction (models/gaze_nerf.py:70-80, 240-243;
// utils/model_utils.py:272-280, 317-318): [d, sin(2^0 d), cos(2^0 d), ..., sin(2^3 d), cos(2^3 d)].  One CTA per 8 rays.
__global__ void __launch_bounds__(kH2) vd_bias_kernel(const float* __restrict__ aux, const float4* __restrict__ ray_dl, int n_rays,
                                                    float* __restrict__ vd_bias) {
  __shared__ float s_pe[8][32];
  const int r0 = blockIdx.x * 8;
  if (threadIdx.x < 8 * 3) {
    const int rl = threadIdx.x / 3, c = threadIdx.x % 3;
    if (r0 + rl < n_rays) {
      const float4 dl = ray_dl[r0 + rl];
      const float x = c == 0 ? dl.x : (c == 1 ? dl.y : dl.z);
      s_pe[rl][c] = x;
      float f = 1.0f;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float sv, cv;
        sincosf(__fmul_rn(x, f), &sv, &cv);
        s_pe[rl][3 + 6 * q + c] = sv;
        s_pe[rl][6 + 6 * q + c] = cv;
        f *= 2.0f;
      }
    }
  }
  __syncthreads();
  const int n = threadIdx.x;
  float w[kVdDims];
#pragma unroll
  for (int j = 0; j < kVdDims; ++j) w[j] = __ldg(aux + kAuxW1v + (size_t)n * 32 + j);
  for (int rl = 0; rl < 8 && r0 + rl < n_rays; ++rl) {
    float acc = 0.0f;
#pragma unroll
    for (int j = 0; j < kVdDims; ++j) acc = fmaf(w[j], s_pe[rl][j], acc);
    vd_bias[(size_t)(r0 + rl) * kH2 + n] = acc;
  }
}

// per-face biases: bias[b] = base + code / appearance contributions (fold i).  One warp per output row: the lanes stride over the
// 181 code (127 appearance) columns of the row, so the weight reads are coalesced; rows without a fold are plain copies.
constexpr int kFoldWarps = 8;
__global__ void __launch_bounds__(kFoldWarps * 32) fold_kernel(const float* __restrict__ aux, const float* __restrict__ shape_ext,
                                                             const float* __restrict__ appea, float* __restrict__ bias) {
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * kFoldWarps + (threadIdx.x >> 5);
  if (i >= kBiasFloats) return;
  const int layer = i / kHidden;
  const int n = i - layer * kHidden;
  const float* w = nullptr;
  const float* c = nullptr;
  int len = 0;
  if (layer == 0 || layer == 5) {
    w = aux + (layer == 0 ? kAuxW0c : kAuxW5c) + (size_t)n * GNRF_SHAPE_EXT_DIMS;
    c = shape_ext + (size_t)b * GNRF_SHAPE_EXT_DIMS;
    len = GNRF_SHAPE_EXT_DIMS;
  } else if (layer == 8 && n < kH2) {
    w = aux + kAuxW1c + (size_t)n * GNRF_APPEA_DIMS;
    c = appea + (size_t)b * GNRF_APPEA_DIMS;
    len = GNRF_APPEA_DIMS;
  }
  float acc = 0.0f;
  for (int k = lane; k < len; k += 32) acc = fmaf(__ldg(w + k), __ldg(c + k), acc);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if (lane == 0) bias[(size_t)b * kBiasFloats + i] = aux[kAuxBaseBias + i] + acc;
}

}  // namespace tc
}  // namespace gnrf

using namespace gnrf;
using namespace gnrf::tc;

extern "C" size_t gnrf_mlp_tc_packed_bytes(void) { return kPackedBytes; }
extern "C" size_t gnrf_mlp_tc_bias_floats(void) { return kBiasFloats; }

extern "C" int gnrf_mlp_tc_pack_vd(const float* const* params, int n_vd, void* packed, gnrf_stream_t stream) {
  GNRF_CHECK_ARG(params && packed);
  GNRF_CHECK_ARG(n_vd == 0 || n_vd == kVdDims);
  GNRF_CHECK_ARG((reinterpret_cast<uintptr_t>(packed) & 127) == 0);
  PackSrc src;
  src.n_vd = n_vd;
  for (int i = 0; i < 12; ++i) {
    src.w[i] = params[2 * i];
    src.b[i] = params[2 * i + 1];
    GNRF_CHECK_ARG(src.w[i] && src.b[i]);
  }
  unsigned char* p = static_cast<unsigned char*>(packed);
  float* aux = reinterpret_cast<float*>(p + kStreamBytes);
  cudaStream_t st = as_stream(stream);
  fuse_head_kernel<<<kL8N, 128, 0, st>>>(src, aux);
  pack_aux_kernel<<<148, 256, 0, st>>>(src, aux);
  pack_stream_kernel<<<kNumStagesPerTile, 256, 0, st>>>(src, aux + kAuxWf, p);
  GNRF_LAUNCH_CHECK();
  count_launches(3);
  return GNRF_OK;
}

extern "C" int gnrf_mlp_tc_pack(const float* const* params, void* packed, gnrf_stream_t stream) {
  return gnrf_mlp_tc_pack_vd(params, 0, packed, stream);
}

extern "C" int gnrf_mlp_tc_vd_bias(const void* packed, const float* ray_dl, int B, int N_r, float* vd_bias, gnrf_stream_t stream) {
  GNRF_CHECK_ARG(packed && ray_dl && vd_bias && B > 0 && N_r > 0);
  const float* aux = reinterpret_cast<const float*>(static_cast<const unsigned char*>(packed) + kStreamBytes);
  vd_bias_kernel<<<ceil_div(B * N_r, 8), kH2, 0, as_stream(stream)>>>(aux, reinterpret_cast<const float4*>(ray_dl), B * N_r, vd_bias);
  GNRF_LAUNCH_CHECK();
  count_launches(1);
  return GNRF_OK;
}

extern "C" int gnrf_mlp_tc_fold(const void* packed, const float* shape_ext, const float* appea, int B, float* bias,
                                gnrf_stream_t stream) {
  GNRF_CHECK_ARG(packed && shape_ext && appea && bias && B > 0);
  const float* aux = reinterpret_cast<const float*>(static_cast<const unsigned char*>(packed) + kStreamBytes);
  dim3 grid(ceil_div(kBiasFloats, kFoldWarps), B);
  fold_kernel<<<grid, kFoldWarps * 32, 0, as_stream(stream)>>>(aux, shape_ext, appea, bias);
  GNRF_LAUNCH_CHECK();
  count_launches(1);
  return GNRF_OK;
}

extern "C" size_t gnrf_mlp_tc_workspace_bytes(int n_branch, int B, int N_r) {
  if (n_branch <= 0 || B <= 0 || N_r <= 0) return 0;
  size_t hc = ((size_t)n_branch * B * N_r * (kH2 + 1) * sizeof(float) + 255) & ~(size_t)255;
  return hc + (size_t)kMaxGrid * 65 * kTile * sizeof(uint32_t);
}

static int mlp_tc_fwd_impl(int n_branch, const void* const* packed, const float* const* bias, const float* const* vd_bias,
                           const float* ray_dl, const float* tvecs, const float* z_edges, int B, int N_r, int N_s, float* const* feat_ray,
                           float* const* bg_alpha, float* const* weights, void* workspace, size_t workspace_bytes, float* dbg_dump,
                           long long* timeline, int cluster_size, gnrf_stream_t stream) {
  GNRF_CHECK_ARG(packed && bias && ray_dl && tvecs && z_edges && feat_ray && bg_alpha && workspace);
  GNRF_CHECK_ARG(n_branch == 1 || n_branch == 2);
  GNRF_CHECK_ARG(B > 0 && N_r > 0 && N_s > 0);
  if (N_s > kTile || (kTile % N_s) != 0 || ((long long)N_r * N_s) % kTile != 0)
    return fail(GNRF_ERR_UNSUPPORTED, "gnrf_mlp_tc_fwd: N_s=%d must divide %d and N_r*N_s=%lld must be a multiple of %d", N_s, kTile,
                (long long)N_r * N_s, kTile);
  if (workspace_bytes < gnrf_mlp_tc_workspace_bytes(n_branch, B, N_r))
    return fail(GNRF_ERR_ARG, "gnrf_mlp_tc_fwd: workspace too small");
  cudaStream_t st = as_stream(stream);
  int n_sm = 0;
  {
    int rc = device_once(kOnceMlpTc, &n_sm, []() -> int {
      GNRF_CUDA(cudaFuncSetAttribute(mlp_tc_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
      GNRF_CUDA(cudaFuncSetAttribute(mlp_tc_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
      GNRF_CUDA(cudaFuncSetAttribute(mlp_tc_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
      GNRF_CUDA(cudaFuncSetAttribute(mlp_tc_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
      return GNRF_OK;
    });
    if (rc != GNRF_OK) return rc;
  }
  FwdArgs a;
  float* ws = static_cast<float*>(workspace);
  for (int i = 0; i < 2; ++i) {
    int j = i < n_branch ? i : 0;
    GNRF_CHECK_ARG(packed[j] && bias[j] && feat_ray[j] && bg_alpha[j]);
    a.br[i].stream = static_cast<const unsigned char*>(packed[j]);
    a.br[i].bias = bias[j];
    a.br[i].vd_bias = vd_bias ? vd_bias[j] : nullptr;
    a.br[i].hc = ws + (size_t)j * B * N_r * (kH2 + 1);
    a.br[i].wsum = a.br[i].hc + (size_t)B * N_r * kH2;
    a.br[i].weights = weights ? weights[j] : nullptr;
  }
  a.ray_dl = reinterpret_cast<const float4*>(ray_dl);
  a.tvecs = tvecs;
  a.z_edges = z_edges;
  a.n_branch = n_branch;
  a.B = B;
  a.N_r = N_r;
  a.N_s = N_s;
  a.tiles_per_face = (int)(((long long)N_r * N_s) / kTile);
  a.n_items = n_branch * B * a.tiles_per_face;
  a.pe_stash = reinterpret_cast<uint32_t*>(static_cast<unsigned char*>(workspace) +
                                           (((size_t)n_branch * B * N_r * (kH2 + 1) * sizeof(float) + 255) & ~(size_t)255));
  a.dbg = dbg_dump;    // optional [10][128][384] fp32 dump of tile 0's activations
  a.prof = timeline;   // optional [4][10][16] int64 clock64 timeline of CTA 0
  // cluster size: CTAs of a cluster share every weight stage through TMA multicast (default 2)
  int csize = (cluster_size == 1) ? 1 : 2;
  while (csize > 1 && ((B * a.tiles_per_face) % csize != 0 || a.n_items / csize < 1)) csize >>= 1;
  int grid = (a.n_items / csize) * csize < n_sm ? (a.n_items / csize) * csize : n_sm;
  if (grid > kMaxGrid) grid = kMaxGrid;
  grid -= grid % csize;
  {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = kSmemBytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)csize;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const bool prof = a.prof != nullptr;
    if (csize == 2) {
      if (prof) GNRF_CUDA(cudaLaunchKernelEx(&cfg, mlp_tc_kernel<2, true>, a));
      else GNRF_CUDA(cudaLaunchKernelEx(&cfg, mlp_tc_kernel<2, false>, a));
    } else {
      if (prof) GNRF_CUDA(cudaLaunchKernelEx(&cfg, mlp_tc_kernel<1, true>, a));
      else GNRF_CUDA(cudaLaunchKernelEx(&cfg, mlp_tc_kernel<1, false>, a));
    }
  }
  GNRF_LAUNCH_CHECK();
  {
    HeadArgs ha;
    for (int i = 0; i < 2; ++i) {
      int j = i < n_branch ? i : 0;
      ha.aux[i] = reinterpret_cast<const float*>(static_cast<const unsigned char*>(packed[j]) + kStreamBytes);
      ha.hc[i] = a.br[j].hc;
      ha.wsum[i] = a.br[j].wsum;
      ha.feat_ray[i] = feat_ray[j];
      ha.bg_alpha[i] = bg_alpha[j];
    }
    dim3 g(ceil_div(N_r, kHeadTM), ceil_div(kFeat, kHeadTN), B * n_branch);   // blockIdx.z = face * n_branch + branch
    rgb_head_kernel<<<g, 256, 0, st>>>(ha, N_r, n_branch);
  }
  GNRF_LAUNCH_CHECK();
  count_launches(2);
  return GNRF_OK;
}

extern "C" int gnrf_mlp_tc_fwd_debug(int n_branch, const void* const* packed, const float* const* bias, const float* ray_dl,
                                     const float* tvecs, const float* z_edges, int B, int N_r, int N_s, float* const* feat_ray,
                                     float* const* bg_alpha, float* const* weights, void* workspace, size_t workspace_bytes,
                                     float* dbg_dump, long long* timeline, int cluster_size, gnrf_stream_t stream) {
  return mlp_tc_fwd_impl(n_branch, packed, bias, nullptr, ray_dl, tvecs, z_edges, B, N_r, N_s, feat_ray, bg_alpha, weights, workspace,
                         workspace_bytes, dbg_dump, timeline, cluster_size, stream);
}

extern "C" int gnrf_mlp_tc_fwd(int n_branch, const void* const* packed, const float* const* bias, const float* ray_dl,
                               const float* tvecs, const float* z_edges, int B, int N_r, int N_s, float* const* feat_ray,
                               float* const* bg_alpha, float* const* weights, void* workspace, size_t workspace_bytes,
                               gnrf_stream_t stream) {
  return mlp_tc_fwd_impl(n_branch, packed, bias, nullptr, ray_dl, tvecs, z_edges, B, N_r, N_s, feat_ray, bg_alpha, weights, workspace,
                         workspace_bytes, nullptr, nullptr, 2, stream);
}

extern "C" int gnrf_mlp_tc_fwd_vd(int n_branch, const void* const* packed, const float* const* bias, const float* const* vd_bias,
                                  const float* ray_dl, const float* tvecs, const float* z_edges, int B, int N_r, int N_s,
                                  float* const* feat_ray, float* const* bg_alpha, float* const* weights, void* workspace,
                                  size_t workspace_bytes, gnrf_stream_t stream) {
  GNRF_CHECK_ARG(vd_bias);
  for (int i = 0; i < n_branch && i < 2; ++i) GNRF_CHECK_ARG(vd_bias[i] != nullptr);
  return mlp_tc_fwd_impl(n_branch, packed, bias, vd_bias, ray_dl, tvecs, z_edges, B, N_r, N_s, feat_ray, bg_alpha, weights, workspace,
                         workspace_bytes, nullptr, nullptr, 2, stream);
}
