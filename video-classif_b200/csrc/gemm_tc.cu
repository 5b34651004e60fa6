// tcgen05 / TMEM / TMA GEMM and implicit-GEMM convolution for sm_100a, with the train-mode
// BatchNorm of a ResNet block folded into its producer and consumer.
//
//   D[M,N] = epi( tf(A)[M,K] * B[N,K]^T )          fp32 accumulate in TMEM, bf16 operands.
//
// A is fed by TMA either from a plain row-major matrix (2-D tiled tensor map: Linear layers,
// 1x1 stride-1 convolutions over NHWC activations, hoisted LSTM gate GEMMs) or straight from an
// NHWC activation tensor through an IM2COL-mode tensor map (3x3 / strided convolutions: one TMA
// per filter tap and 64-channel slab, zero padding and stride handled by the copy engine), so no
// im2col matrix ever exists in HBM.  B (the weights, [N,K] K-major) always comes from a 2-D map.
//
// BatchNorm fusion (train mode needs the statistics of the WHOLE conv output before anything can be
// normalised, so it is split across kernels instead of costing extra passes over the activations):
//   * producer epilogue: per-column sum / sum-of-squares of the raw conv output (col_sum/col_sumsq);
//     the last CTA to finish turns them into per-channel scale/shift (+ running-stat update): BnFinal
//   * consumer prologue ("A transform"): 4 extra warps rewrite each A tile in shared memory as
//     relu(a * scale[c] + shift[c]) between the TMA arrival and the MMA, so the normalised tensor is
//     never materialised (zero-padding taps stay zero)
//   * block output: the conv3 GEMM runs twice -- a statistics-only pass (store = 0), then a pass whose
//     epilogue applies BN3, adds the shortcut (identity, or the raw down-sample conv with its own
//     BN) and ReLUs -- instead of writing the raw tensor and re-reading it.
//
// The kernel is specialised at compile time by epilogue MODE and by the presence of the A transform, so
// that every instantiation is a few hundred straight-line SASS instructions per role: the first
// monolithic version (8.6k instructions, every option a run-time branch) was instruction-fetch bound in
// the epilogue warps (ncu: stall_no_inst on most epilogue instructions).
//   EPI_GENERIC  bias / bias2 / ReLU, fp32 or bf16 output, optional statistics (Linear layers, tests)
//   EPI_BF16     raw bf16 conv output through TMA bulk stores (+ optional statistics, finalisation)
//   EPI_STATS    statistics-only pass (nothing stored)
//   EPI_POST     BatchNorm of the output + shortcut (+ its BatchNorm) + ReLU, bf16 through TMA stores
//
// Kernel anatomy: persistent CTAs (one per SM), 10 warps (+4 with the A transform):
//   warp 0        TMA producer           (smem ring of kStages {A 128x64, B BNx64} bf16 tiles, SW128)
//   warp 1        MMA issuer + TMEM owner (tcgen05.mma cta_group::1, M=128, N=BN, K=16 per instr)
//   warps 2..9    epilogue               (tcgen05.ld 32x32b -> bias / BN / residual / ReLU -> bf16 ->
//                                         swizzled staging tile -> TMA bulk store; column statistics)
//   warps 10..13  A transform            (idle unless at.scale is set)
// TMEM holds two BN-column accumulators so the epilogue of tile i overlaps the MMAs of tile i+1.
#include "tc_ptx.cuh"

#include <cudaTypedefs.h>
#include <mutex>
#include <stdlib.h>

namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 B = one SWIZZLE_128B row
constexpr int UMMA_K = 16;
constexpr int kEpiWarps = 8;                      // two warps per TMEM lane quarter
constexpr int kTfWarps = 4;                       // A-transform warps
constexpr int kEpiThread0 = 64;
constexpr int kTfThread0 = 64 + 32 * kEpiWarps;
constexpr int kThreadsTf = kTfThread0 + 32 * kTfWarps;   // with the A transform
constexpr int kThreadsNoTf = kTfThread0;                 // without: the 4 transform warps do not exist

enum EpiMode { EPI_GENERIC = 0, EPI_BF16 = 1, EPI_STATS = 2, EPI_POST = 3 };

using namespace tc;

struct ConvGeom {   // im2col producer geometry (all zero for plain GEMM)
  int is_conv;
  int P, Q;         // output height / width
  int S;            // filter width (taps per filter row)
  int c_slabs;      // C / 64
  int stride;
  int lower_w, lower_h;   // = -pad
  int H, W;         // input height / width (padding mask of the A transform)
};

struct ATransform {      // a' = relu?(a * scale[c] + shift[c]) applied to the A operand in shared memory
  const float* scale;    // [K channels] or null (no transform)
  const float* shift;
  int relu;
};

struct BnFinal {         // BatchNorm finalisation by the last CTA: (col_sum, col_sumsq) -> scale/shift
  float* scale;          // [N] or null (no finalisation)
  float* shift;
  const float* gamma;
  const float* beta;
  float* running_mean;   // may be null
  float* running_var;
  unsigned int* counter; // zeroed by the caller before the launch
  float inv_count, unbias, eps, momentum;
};

struct EpiParams {
  void* D;
  long ldd;          // elements
  const float* bias; // [N] or null
  const float* bias2; // [N] or null (second bias vector, e.g. LSTM b_hh)
  float* col_sum;    // [N] or null  (atomicAdd) -- statistics of the raw (pre-BN) output
  float* col_sumsq;  // [N] or null
  int out_bf16;      // 1: bf16 out, 0: fp32 out
  int relu;
  int tma_store;     // bf16 output leaves through smem staging + TMA bulk stores (tmap_d valid)
  int store;         // 0: statistics-only pass, D is never written
  int k_splits;      // > 1 (plain GEMM, fp32 output, EPI_GENERIC): gridDim.y CTAs share a tile's K range and ADD into a zeroed D
  const float* o_scale;   // [N] or null: BatchNorm of the output applied in the epilogue
  const float* o_shift;
  const bf16* res;        // [M][ldres] or null: shortcut added before the ReLU
  long ldres;
  const float* r_scale;   // [N] or null: BatchNorm of the shortcut (down-sample branch)
  const float* r_shift;
  BnFinal fin;
};

constexpr int kStgBytes = 32 * 64;  // epilogue staging tile: 32 rows x 32 bf16 (64-byte rows, SWIZZLE_64B)

constexpr int kResSlots = 4;         // EPI_POST: per-warp ring of shortcut / output tiles (power of two)

template <int BN, int MODE, bool DEEP = false, bool PAIR = false>
struct SmemLayout {
  static constexpr bool kPost = MODE == EPI_POST;
  // PAIR (cta_group::2): each CTA of the pair stages its own 128 rows of A and HALF of the B tile
  static constexpr int kStageBytes = (BM * BK + (PAIR ? BN / 2 : BN) * BK) * 2;
  // EPI_POST trades operand stages for the 64 KB shortcut ring (its GEMMs are short-K and memory bound).
  // DEEP (BN = 256, long K, tensor bound): a 4th operand stage paid for with the second staging tile -- three
  // 48 KB stages cover ~0.8 us of TMA latency, less than the latency of an L2 hit under load
  static constexpr int kStages = PAIR ? (kPost ? 4 : 5)
                                 : kPost ? (BN >= 256 ? 3 : (BN >= 128 ? 4 : 5))
                                         : (BN >= 256 ? (DEEP ? 4 : 3) : (BN >= 128 ? 5 : (BN >= 64 ? 7 : 8)));
  static constexpr int kStgBufs = DEEP ? 1 : 2;          // output staging tiles per epilogue warp (not EPI_POST)
  static constexpr int kTileBytes = kStageBytes * kStages;
  static constexpr int kBarOffset = kTileBytes;
  static constexpr int kNumBars = 3 * kStages + 4 + kEpiWarps * kResSlots;
  // [4 quarters x 2 half-warps][2][BN] column statistics at flush time; the SAME bytes hold the per-column
  // {o_scale, o_shift, r_scale, r_shift}[BN] of an output pass (statistics and output BN never mix)
  static constexpr int kStatOffset = (kBarOffset + kNumBars * 8 + 16 + 15) / 16 * 16;
  static constexpr int kScratchOffset = kStatOffset + 8 * 2 * BN * 4;
  static constexpr int kScratchOffset1k = (kScratchOffset + 1023) / 1024 * 1024;        // tiles 1024-aligned
  // non-POST: kEpiWarps x kStgBufs staging tiles; POST: kEpiWarps x kResSlots shortcut-in / output-out tiles
  static constexpr int kScratchBytes = kEpiWarps * (kPost ? kResSlots : kStgBufs) * kStgBytes;
  static constexpr int kTotal = kScratchOffset1k + kScratchBytes + 1024;   // +1024 alignment slack
  static_assert(kTotal <= 227 * 1024, "shared memory budget");
};

__device__ __forceinline__ uint32_t pack_bf16x2_relu(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}

template <int BN, int MODE, bool TF, bool CL = false, bool DEEP = false, bool PAIR = false>
__global__ void __launch_bounds__(TF ? kThreadsTf : kThreadsNoTf, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const __grid_constant__ CUtensorMap tmap_d, const __grid_constant__ CUtensorMap tmap_r, int M, int N,
               int K, ConvGeom g, ATransform at, EpiParams ep) {
  static_assert(!PAIR || (!TF && !CL && BN == 256 && MODE != EPI_GENERIC), "CTA pairs: 256-column conv epilogues without the A transform");
  using L = SmemLayout<BN, MODE, DEEP, PAIR>;
  constexpr int kStages = L::kStages;
  constexpr uint32_t kTmemCols = 2 * BN;  // two accumulators (32 <= cols <= 512, power of two)
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte alignment (SWIZZLE_128B) computed as an OFFSET so that every derived pointer keeps its
  // shared-memory provenance (integer round-tripping made the compiler emit generic LD/ST/atomics).
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tf_bar = empty_bar + kStages;
  uint64_t* tfull_bar = tf_bar + kStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* res_bar = tempty_bar + 2;            // kResSlots per epilogue warp (EPI_POST)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + kEpiWarps * kResSlots);
  // Column statistics (sum, sum of squares of the stored values) live in REGISTERS of the epilogue
  // warps (each warp owns fixed 32-column chunks of the n-block) over all consecutive tiles of one
  // n-block; they are combined through this buffer ([4 quarters][2][BN], one writer per slot) and
  // flushed to global memory only when the n-block changes (tiles are ordered m-fastest).
  float* stat_s = reinterpret_cast<float*>(smem + L::kStatOffset);
  uint8_t* staging_s = smem + L::kScratchOffset1k;      // EPI_POST: the shortcut / output ring lives here
  constexpr int kThreads = TF ? kThreadsTf : kThreadsNoTf;
  constexpr bool kGeneric = MODE == EPI_GENERIC;
  constexpr bool kPost = MODE == EPI_POST;
  constexpr bool kStore = MODE != EPI_STATS;
  const bool want_stats = MODE == EPI_STATS || (MODE != EPI_POST && ep.col_sum != nullptr);
  for (int i = threadIdx.x; i < 16 * BN; i += kThreads) stat_s[i] = 0.f;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // CL: clusters of two CTAs work on the tile pair (2p, 2p+1) -- same weight columns, adjacent rows (tiles are
  // ordered m-fastest and the grid is even, so blockIdx.x parity = tile parity = cluster rank).  Each CTA fetches HALF
  // of the weight tile and TMA-multicasts it into both CTAs' stages: the L2 -> SM traffic of the B operand, the
  // larger one at BN = 256, halves.  A slot is refilled only after BOTH CTAs' MMAs released it (empty count 2).
  // The row-block count is padded to even; an all-out-of-range tile loads zeros and stores nothing.
  // PAIR: a pair of CTAs (cluster ranks 0 / 1) owns the 256-row tile {2 mt, 2 mt + 1} x one n-block; the tile index space
  // is (row pairs x n-blocks) walked by pair id.  An odd row-block count leaves an all-out-of-range block (zeros in, nothing out)
  const int pair_rank = PAIR ? (int)cluster_ctarank() : 0;
  const int m_blocks = CL ? (((M + BM - 1) / BM + 1) & ~1) : (PAIR ? ((M + BM - 1) / BM + 1) / 2 : (M + BM - 1) / BM);
  const int n_blocks = (N + BN - 1) / BN;
  const int num_tiles = m_blocks * n_blocks;
  const int tile0 = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int tile_step = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  auto row_block = [&](int mt) { return PAIR ? 2 * mt + pair_rank : mt; };    // tile-space row index -> 128-row block
  // split-K (skinny products such as the LSTM gate GEMM [B*T,16384] x [16384,4H]: 20 output tiles would leave 128 SMs idle
  // on a 47 MB operand stream): blockIdx.y owns k-blocks [kb0, kb0 + k_blocks) and its epilogue adds into D
  const int k_blocks_all = (K + BK - 1) / BK;
  const int kb_per = MODE == EPI_GENERIC && gridDim.y > 1 ? (k_blocks_all + (int)gridDim.y - 1) / (int)gridDim.y : k_blocks_all;
  const int kb0 = MODE == EPI_GENERIC ? (int)blockIdx.y * kb_per : 0;
  const int k_blocks = k_blocks_all - kb0 < kb_per ? k_blocks_all - kb0 : kb_per;     // >= 1 (the host sizes gridDim.y)
  const bool split_k = MODE == EPI_GENERIC && gridDim.y > 1;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], CL ? 2 : 1);     // (PAIR: one multicast commit by the leader frees the slot in both CTAs)
      mbar_init(&tf_bar[i], 32 * kTfWarps);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], PAIR ? 2 * kEpiWarps : kEpiWarps);     // PAIR: the epilogue warps of BOTH CTAs release it
    }
    for (int i = 0; i < kEpiWarps * kResSlots; ++i) mbar_init(&res_bar[i], 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    if (PAIR) {
      tc_alloc_pair(tmem_slot, kTmemCols);
      tc_relinquish_pair();
    } else {
      tc_alloc(tmem_slot, kTmemCols);
      tc_relinquish();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CL || PAIR) cluster_sync();  // the peer's barriers exist before any multicast / remote arrive can reach them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  grid_dependency_wait();          // PDL: the set-up above overlapped the previous kernel's tail; its results are visible now
  grid_launch_dependents();

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = tile0; tile < num_tiles; tile += tile_step) {
        const int n_blk = tile / m_blocks;           // m-fastest: concurrent CTAs share one weight tile
        const int m_blk = row_block(tile - n_blk * m_blocks);
        int cn = 0, cw = 0, ch = 0;
        if (g.is_conv) {
          const int m0 = m_blk * BM;
          const int pq = g.P * g.Q;
          cn = m0 / pq;
          const int rem = m0 - cn * pq;
          const int p = rem / g.Q;
          const int q = rem - p * g.Q;
          cw = g.lower_w + q * g.stride;
          ch = g.lower_h + p * g.stride;
        }
        int tap = 0, slab = 0;                       // k-block -> (filter tap, 64-channel slab) without divisions
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * L::kStageBytes;
          uint8_t* sb = sa + BM * BK * 2;
          if (PAIR) {
            // the LEADER's barrier collects the bytes of both CTAs (each: its A rows + its half of B)
            if (pair_rank == 0) mbar_expect_tx(&full_bar[stage], 2 * L::kStageBytes);
            if (g.is_conv) {
              const int r = tap / g.S;
              const int s = tap - r * g.S;
              tma_load_im2col_4d_pair(sa, &tmap_a, &full_bar[stage], slab * BK, cw, ch, cn, (uint16_t)s, (uint16_t)r);
              if (++slab == g.c_slabs) {
                slab = 0;
                ++tap;
              }
            } else {
              tma_load_2d_pair(sa, &tmap_a, &full_bar[stage], (kb0 + kb) * BK, m_blk * BM);
            }
            tma_load_2d_pair(sb, &tmap_b, &full_bar[stage], (kb0 + kb) * BK, n_blk * BN + pair_rank * (BN / 2));
            if (++stage == kStages) {
              stage = 0;
              phase ^= 1;
            }
            continue;
          }
          mbar_expect_tx(&full_bar[stage], L::kStageBytes);
          if (g.is_conv) {
            const int r = tap / g.S;
            const int s = tap - r * g.S;
            tma_load_im2col_4d(sa, &tmap_a, &full_bar[stage], slab * BK, cw, ch, cn, (uint16_t)s, (uint16_t)r);
            if (++slab == g.c_slabs) {
              slab = 0;
              ++tap;
            }
          } else {
            tma_load_2d(sa, &tmap_a, &full_bar[stage], (kb0 + kb) * BK, m_blk * BM);
          }
          if (CL) {
            const int rank = (int)cluster_ctarank();
            tma_load_2d_multicast(sb + rank * (BN / 2) * BK * 2, &tmap_b, &full_bar[stage], (kb0 + kb) * BK,
                                  n_blk * BN + rank * (BN / 2), (uint16_t)3);
          } else {
            tma_load_2d(sb, &tmap_b, &full_bar[stage], (kb0 + kb) * BK, n_blk * BN);
          }
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    constexpr uint32_t idesc = make_idesc(PAIR ? 2 * BM : BM, BN);
    uint64_t* ready_bar = TF ? tf_bar : full_bar;   // with a transform the MMA waits for the rewritten tile
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    if (!PAIR || pair_rank == 0)                    // PAIR: only the leader issues (its peer's warp 1 owns the TMEM allocation)
    for (int tile = tile0; tile < num_tiles; tile += tile_step, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
      for (int kb = 0; kb < k_blocks; ++kb) {
        mbar_wait(&ready_bar[stage], phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = smem_u32(smem + stage * L::kStageBytes);
          const uint32_t sb = sa + BM * BK * 2;
          const uint64_t da = make_sw128_desc(sa);
          const uint64_t db = make_sw128_desc(sb);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // +32 B per K=16 step inside the 128 B swizzle row (encoded >>4 -> +2)
            if (PAIR) tc_mma_bf16_pair(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) != 0);
            else tc_mma_bf16(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) != 0);
          }
          if (PAIR) {
            tc_commit_pair(&empty_bar[stage], (uint16_t)3);                          // frees the slot in BOTH CTAs
            if (kb == k_blocks - 1) tc_commit_pair(&tfull_bar[acc], (uint16_t)3);    // both epilogues: accumulator complete
          } else {
          if (CL) tc_commit_multicast(&empty_bar[stage], (uint16_t)3);   // frees the slot in BOTH CTAs
          else tc_commit(&empty_bar[stage]);                  // frees the smem slot when the MMAs retire
          if (kb == k_blocks - 1) tc_commit(&tfull_bar[acc]);  // accumulator complete -> epilogue
          }
        }
        __syncwarp();
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp < 2 + kEpiWarps) {
    // =========================== epilogue (warps 2..9) ===========================
    constexpr int kChunks = BN / 32;
    constexpr int kMyChunks = (kChunks + 1) / 2;   // chunks per warp: ch = half + 2 j
    const int quarter = warp & 3;          // TMEM lane quarter this warp may read (hardware rule: warp % 4)
    const int half = (warp - 2) >> 2;      // the two warps of a quarter take alternate 32-column chunks
    const int epi_tid = threadIdx.x - kEpiThread0;
    const int sw = (lane >> 1) & 3;        // SWIZZLE_64B phase of this lane's row in a 32x32 bf16 tile
    int it = 0;
    int cur_nblk = -1;
    if constexpr (kPost) {
      // ---- BatchNorm of the output + shortcut (+ its BatchNorm) + ReLU on the fp32 accumulators.
      // Each warp owns a ring of kResSlots 32x32 bf16 tiles: the shortcut tile of chunk-iteration i arrives
      // by TMA kResSlots-1 iterations ahead (48 KB of shortcut loads in flight per SM: with one tile of
      // look-ahead the kernel was latency-bound at ~1.2 TB/s), is combined IN PLACE with the accumulators
      // and leaves by a TMA bulk store from the same tile.
      constexpr int R = kResSlots;
      uint8_t* ring = staging_s + (warp - 2) * R * kStgBytes;
      uint64_t* rbar = res_bar + (warp - 2) * R;
      const bool has_res = ep.res != nullptr;
      const bool has_obn = ep.o_scale != nullptr;
      const bool has_rbn = ep.r_scale != nullptr;
      const bool relu = ep.relu != 0;
      float* ss_s = stat_s;                // [4][BN]: o_scale, o_shift, r_scale, r_shift of the current n-block
      const int my_tiles = (tile0 < num_tiles) ? (num_tiles - tile0 + tile_step - 1) / tile_step : 0;
      const int total_iters = my_tiles * kMyChunks;
      auto issue_res = [&](int i) {        // chunk-iteration i -> (tile, chunk) -> slot i % R
        const int tile = tile0 + (i / kMyChunks) * tile_step;
        const int j = i - (i / kMyChunks) * kMyChunks;
        const int n_blk = tile / m_blocks;
        const int m_blk = row_block(tile - n_blk * m_blocks);
        uint64_t* bar = &rbar[i & (R - 1)];
        mbar_expect_tx(bar, kStgBytes);
        tma_load_2d(ring + (i & (R - 1)) * kStgBytes, &tmap_r, bar, n_blk * BN + (half + 2 * j) * 32,
                    m_blk * BM + quarter * 32);
      };
      if (has_res && lane == 0)
        for (int i = 0; i < R - 1 && i < total_iters; ++i) issue_res(i);
      for (int tile = tile0; tile < num_tiles; tile += tile_step, ++it) {
        const int n_blk = tile / m_blocks;
        const int m_blk = row_block(tile - n_blk * m_blocks);
        if (n_blk != cur_nblk) {
          // ---- per-column BatchNorm coefficients of the new n-block -> shared memory
          asm volatile("bar.sync 1, 256;" ::: "memory");       // everyone is done with the previous block's values
          for (int c = epi_tid; c < BN; c += 32 * kEpiWarps) {
            const int col = n_blk * BN + c;
            const bool ok = col < N;
            ss_s[c] = (has_obn && ok) ? __ldg(ep.o_scale + col) : 1.f;
            ss_s[BN + c] = (has_obn && ok) ? __ldg(ep.o_shift + col) : 0.f;
            ss_s[2 * BN + c] = (has_rbn && ok) ? __ldg(ep.r_scale + col) : 1.f;
            ss_s[3 * BN + c] = (has_rbn && ok) ? __ldg(ep.r_shift + col) : 0.f;
          }
          asm volatile("bar.sync 1, 256;" ::: "memory");
          cur_nblk = n_blk;
        }
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tfull_bar[acc], acc_phase);
        tc_fence_after();
#pragma unroll 1
        for (int j = 0; j < kMyChunks; ++j) {
          const int ch = half + 2 * j;
          const int col0 = n_blk * BN + ch * 32;
          const int i = it * kMyChunks + j;
          uint8_t* slot = ring + (i & (R - 1)) * kStgBytes;
          if (col0 < N) {   // warp-uniform (always true with a shortcut: the host requires N % BN == 0)
            uint32_t raw[32];
            tc_ld32(tmem_base + (uint32_t)(acc * BN + ch * 32) + ((uint32_t)(quarter * 32) << 16), raw);
            uint4 rres[4];
            if (has_res) {   // own row (64 B) of the swizzled shortcut tile
              mbar_wait(&rbar[i & (R - 1)], (uint32_t)(i / R) & 1u);
#pragma unroll
              for (int c = 0; c < 4; ++c) rres[c] = *reinterpret_cast<const uint4*>(slot + lane * 64 + ((c ^ sw) << 4));
            } else {         // the slot's previous bulk store (R iterations ago) must have been read out
              if (lane == 0) bulk_wait_read<R - 1>();
              __syncwarp();
            }
            tc_wait_ld();
            float v[32];
#pragma unroll
            for (int jj = 0; jj < 32; ++jj) v[jj] = __uint_as_float(raw[jj]);
            const float* sc_p = ss_s + ch * 32;
            if (has_obn) {
#pragma unroll
              for (int jj = 0; jj < 32; jj += 4) {
                const float4 sc = *reinterpret_cast<const float4*>(sc_p + jj);
                const float4 sh = *reinterpret_cast<const float4*>(sc_p + BN + jj);
                v[jj] = fmaf(v[jj], sc.x, sh.x);
                v[jj + 1] = fmaf(v[jj + 1], sc.y, sh.y);
                v[jj + 2] = fmaf(v[jj + 2], sc.z, sh.z);
                v[jj + 3] = fmaf(v[jj + 3], sc.w, sh.w);
              }
            }
            if (has_res) {
              const uint32_t* rw = reinterpret_cast<const uint32_t*>(rres);
              if (has_rbn) {
#pragma unroll
                for (int jj = 0; jj < 32; jj += 4) {
                  const float4 sc = *reinterpret_cast<const float4*>(sc_p + 2 * BN + jj);
                  const float4 sh = *reinterpret_cast<const float4*>(sc_p + 3 * BN + jj);
                  const uint32_t w0 = rw[jj >> 1], w1 = rw[(jj >> 1) + 1];
                  v[jj] += fmaf(__uint_as_float(w0 << 16), sc.x, sh.x);
                  v[jj + 1] += fmaf(__uint_as_float(w0 & 0xffff0000u), sc.y, sh.y);
                  v[jj + 2] += fmaf(__uint_as_float(w1 << 16), sc.z, sh.z);
                  v[jj + 3] += fmaf(__uint_as_float(w1 & 0xffff0000u), sc.w, sh.w);
                }
              } else {
#pragma unroll
                for (int jj = 0; jj < 16; ++jj) {
                  v[2 * jj] += __uint_as_float(rw[jj] << 16);
                  v[2 * jj + 1] += __uint_as_float(rw[jj] & 0xffff0000u);
                }
              }
            }
            uint32_t pk[16];
            if (relu) {
#pragma unroll
              for (int jj = 0; jj < 16; ++jj) pk[jj] = pack_bf16x2_relu(v[2 * jj], v[2 * jj + 1]);
            } else {
#pragma unroll
              for (int jj = 0; jj < 16; ++jj) pk[jj] = pack_bf16x2(v[2 * jj], v[2 * jj + 1]);
            }
#pragma unroll
            for (int c = 0; c < 4; ++c)
              *reinterpret_cast<uint4*>(slot + lane * 64 + ((c ^ sw) << 4)) =
                  make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&tmap_d, slot, col0, m_blk * BM + quarter * 32);
              bulk_commit();
              if (has_res && i + R - 1 < total_iters) {
                bulk_wait_read<1>();       // the store of iteration i-1 has left its slot: refill it
                issue_res(i + R - 1);
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (PAIR) mbar_arrive_leader(&tempty_bar[acc]);
          else mbar_arrive(&tempty_bar[acc]);
        }
      }
      if (lane == 0) bulk_wait_all();
    } else {
    uint8_t* stg_base = staging_s + (warp - 2) * L::kStgBufs * kStgBytes;
    // options an instantiation does not have fold to constants (the code behind them disappears)
    const bool out_bf16 = kGeneric ? ep.out_bf16 != 0 : true;
    const bool relu = kGeneric ? ep.relu != 0 : false;
    const bool tma_store = kGeneric ? ep.tma_store != 0 : kStore;
    const bool stats_bf16 = want_stats && out_bf16;
    int stg_buf = 0;
    // statistics: lane (w = lane & 15, hf = lane >> 4) sums columns 2w, 2w+1 over rows 16 hf .. 16 hf + 15;
    // the per-warp partial sums accumulate in stat_s (one writer per slot) until the n-block changes
    const int sw_w = lane & 15, sw_hf = lane >> 4;
    // staging reads: row r = i (hf = 0) or 16 + (i ^ 1) (hf = 1) so the two half-warps never share a
    // bank; word address = r*64 + (((w >> 2) ^ ((r >> 1) & 3)) << 4) + (w & 3) * 4
    uint32_t rd_off[2][4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t col = (uint32_t)((((sw_w >> 2) ^ k) << 4) + (sw_w & 3) * 4) + (uint32_t)(sw_hf * 1024);
      rd_off[0][k] = col + (uint32_t)(sw_hf * 64);    // even i: row i + hf
      rd_off[1][k] = col - (uint32_t)(sw_hf * 64);    // odd i:  row i - hf
    }

    for (int tile = tile0; tile < num_tiles; tile += tile_step, ++it) {
      const int n_blk = tile / m_blocks;
      const int m_blk = row_block(tile - n_blk * m_blocks);
      if (n_blk != cur_nblk) {
        if (want_stats && cur_nblk >= 0) {
          // ---- flush the finished n-block: stat_s (4 quarter copies) -> global atomics, then clear
          asm volatile("bar.sync 1, 256;" ::: "memory");
          for (int c = epi_tid; c < BN; c += 32 * kEpiWarps) {
            const int col = cur_nblk * BN + c;
            if (col < N) {
              float a1 = 0.f, a2 = 0.f;
#pragma unroll
              for (int qq = 0; qq < 8; ++qq) {
                a1 += stat_s[qq * 2 * BN + c];
                a2 += stat_s[qq * 2 * BN + BN + c];
              }
              atomicAdd(ep.col_sum + col, a1);
              atomicAdd(ep.col_sumsq + col, a2);
            }
#pragma unroll
            for (int qq = 0; qq < 16; ++qq) stat_s[qq * BN + c] = 0.f;
          }
          asm volatile("bar.sync 1, 256;" ::: "memory");
        }
        cur_nblk = n_blk;
      }
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int row = m_blk * BM + quarter * 32 + lane;
      const bool row_ok = row < M;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      // TMEM reads are software-pipelined: the tcgen05.ld of chunk j+1 is in flight while chunk j is converted,
      // reduced and staged, and the accumulator is handed back to the MMA warp as soon as the last read has landed
      auto chunk_ok = [&](int j) { return j < kMyChunks && half + 2 * j < kChunks && n_blk * BN + (half + 2 * j) * 32 < N; };
      auto issue = [&](int j, uint32_t (&raw)[32]) {
        tc_ld32(tmem_base + (uint32_t)(acc * BN + (half + 2 * j) * 32) + ((uint32_t)(quarter * 32) << 16), raw);
      };
      auto release_acc = [&]() {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (PAIR) mbar_arrive_leader(&tempty_bar[acc]);
          else mbar_arrive(&tempty_bar[acc]);
        }
      };
      auto process = [&](int j, const uint32_t (&raw)[32]) {
        const int ch = half + 2 * j;
        const int col0 = n_blk * BN + ch * 32;
        {
          const bool full_chunk = (col0 + 32 <= N);
          float v[32];
#pragma unroll
          for (int jj = 0; jj < 32; ++jj) v[jj] = __uint_as_float(raw[jj]);
          // every option below is a WARP-UNIFORM branch around its own loop
          if (kGeneric && ep.bias != nullptr && kb0 == 0) {
            if (full_chunk) {
#pragma unroll
              for (int jj = 0; jj < 32; jj += 4) {
                const float4 bv = __ldg(reinterpret_cast<const float4*>(ep.bias + col0 + jj));
                v[jj] += bv.x; v[jj + 1] += bv.y; v[jj + 2] += bv.z; v[jj + 3] += bv.w;
              }
            } else {
              for (int jj = 0; jj < 32; ++jj)
                if (col0 + jj < N) v[jj] += __ldg(ep.bias + col0 + jj);
            }
          }
          if (kGeneric && ep.bias2 != nullptr && kb0 == 0) {
            for (int jj = 0; jj < 32; ++jj)
              if (col0 + jj < N) v[jj] += __ldg(ep.bias2 + col0 + jj);
          }
          // statistics are taken on the bf16-rounded value as stored
          uint32_t pk[16];
          if (out_bf16) {
            if (relu) {
#pragma unroll
              for (int jj = 0; jj < 16; ++jj) pk[jj] = pack_bf16x2_relu(v[2 * jj], v[2 * jj + 1]);
            } else {
#pragma unroll
              for (int jj = 0; jj < 16; ++jj) pk[jj] = pack_bf16x2(v[2 * jj], v[2 * jj + 1]);
            }
          }
          uint8_t* stg = stg_base + stg_buf * kStgBytes;
          if (stats_bf16) {
            // 32x32 bf16 chunk -> 64-byte-swizzled staging tile (lane = row, 4 x STS.128, conflict-free);
            // column sums are read straight back from the tile.
            if (tma_store && lane == 0) bulk_wait_read<L::kStgBufs - 1>();   // tile free again?
            __syncwarp();
#pragma unroll
            for (int c = 0; c < 4; ++c)
              *reinterpret_cast<uint4*>(stg + lane * 64 + ((c ^ sw) << 4)) =
                  row_ok ? make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]) : make_uint4(0u, 0u, 0u, 0u);
            __syncwarp();
            float2 s1a = make_float2(0.f, 0.f), s1b = s1a, s2a = s1a, s2b = s1a;
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
              const uint32_t u0 = *reinterpret_cast<const uint32_t*>(stg + rd_off[0][(i >> 1) & 3] + i * 64);
              const uint32_t u1 = *reinterpret_cast<const uint32_t*>(stg + rd_off[1][(i >> 1) & 3] + (i + 1) * 64);
              const float2 x0 = make_float2(__uint_as_float(u0 << 16), __uint_as_float(u0 & 0xffff0000u));
              const float2 x1 = make_float2(__uint_as_float(u1 << 16), __uint_as_float(u1 & 0xffff0000u));
              s1a = __fadd2_rn(s1a, x0);
              s1b = __fadd2_rn(s1b, x1);
              s2a = __ffma2_rn(x0, x0, s2a);
              s2b = __ffma2_rn(x1, x1, s2b);
            }
            // each half-warp owns its own copy of the statistics (no shuffle, no atomics: one writer per slot)
            float2* st = reinterpret_cast<float2*>(stat_s + (quarter * 2 + sw_hf) * 2 * BN + ch * 32 + 2 * sw_w);
            float2* st2 = reinterpret_cast<float2*>(stat_s + (quarter * 2 + sw_hf) * 2 * BN + BN + ch * 32 + 2 * sw_w);
            *st = __fadd2_rn(*st, __fadd2_rn(s1a, s1b));
            *st2 = __fadd2_rn(*st2, __fadd2_rn(s2a, s2b));
          }
          if (kStore) {
            if (out_bf16) {
              if (tma_store) {
                // staging tile -> one TMA bulk store in full 64-byte row segments (no LSU, no registers)
                if (!stats_bf16) {
                  if (lane == 0) bulk_wait_read<L::kStgBufs - 1>();
                  __syncwarp();
#pragma unroll
                  for (int c = 0; c < 4; ++c)
                    *reinterpret_cast<uint4*>(stg + lane * 64 + ((c ^ sw) << 4)) =
                        make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                  tma_store_2d(&tmap_d, stg, col0, m_blk * BM + quarter * 32);
                  bulk_commit();
                }
                stg_buf = (stg_buf + 1) % L::kStgBufs;
              } else if (kGeneric && row_ok) {
                bf16* dp = reinterpret_cast<bf16*>(ep.D) + (long)row * ep.ldd + col0;
                if (full_chunk && ((reinterpret_cast<uintptr_t>(dp) & 31) == 0)) {
                  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dp), "r"(pk[0]), "r"(pk[1]),
                               "r"(pk[2]), "r"(pk[3]), "r"(pk[4]), "r"(pk[5]), "r"(pk[6]), "r"(pk[7])
                               : "memory");
                  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dp + 16), "r"(pk[8]), "r"(pk[9]),
                               "r"(pk[10]), "r"(pk[11]), "r"(pk[12]), "r"(pk[13]), "r"(pk[14]), "r"(pk[15])
                               : "memory");
                } else if (full_chunk && ((reinterpret_cast<uintptr_t>(dp) & 15) == 0)) {
#pragma unroll
                  for (int jj = 0; jj < 16; jj += 4)
                    *reinterpret_cast<uint4*>(dp + 2 * jj) = make_uint4(pk[jj], pk[jj + 1], pk[jj + 2], pk[jj + 3]);
                } else {
                  for (int jj = 0; jj < 32; ++jj)
                    if (col0 + jj < N) dp[jj] = __ushort_as_bfloat16((unsigned short)(pk[jj >> 1] >> ((jj & 1) * 16)));
                }
              }
            } else if (kGeneric && row_ok) {
              if (relu) {
#pragma unroll
                for (int jj = 0; jj < 32; ++jj) v[jj] = fmaxf(v[jj], 0.0f);
              }
              float* dp = reinterpret_cast<float*>(ep.D) + (long)row * ep.ldd + col0;
              if (split_k) {                    // partial sum of this K range: vector reductions into the zeroed output
                if (full_chunk && ((reinterpret_cast<uintptr_t>(dp) & 15) == 0)) {
#pragma unroll
                  for (int jj = 0; jj < 32; jj += 4)
                    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dp + jj), "f"(v[jj]), "f"(v[jj + 1]),
                                 "f"(v[jj + 2]), "f"(v[jj + 3])
                                 : "memory");
                } else {
                  for (int jj = 0; jj < 32; ++jj)
                    if (col0 + jj < N) atomicAdd(dp + jj, v[jj]);
                }
              } else if (full_chunk && ((reinterpret_cast<uintptr_t>(dp) & 31) == 0)) {
#pragma unroll
                for (int jj = 0; jj < 32; jj += 8)
                  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dp + jj), "f"(v[jj]), "f"(v[jj + 1]),
                               "f"(v[jj + 2]), "f"(v[jj + 3]), "f"(v[jj + 4]), "f"(v[jj + 5]), "f"(v[jj + 6]), "f"(v[jj + 7])
                               : "memory");
              } else if (full_chunk && ((reinterpret_cast<uintptr_t>(dp) & 15) == 0)) {
#pragma unroll
                for (int jj = 0; jj < 32; jj += 4)
                  *reinterpret_cast<float4*>(dp + jj) = make_float4(v[jj], v[jj + 1], v[jj + 2], v[jj + 3]);
              } else {
                for (int jj = 0; jj < 32; ++jj)
                  if (col0 + jj < N) dp[jj] = v[jj];
              }
            }
          }
          if (kGeneric && want_stats && !out_bf16) {
            // fp32 output (statistics of the value as stored, ReLU included): column sums over this warp's
            // 32 rows by butterfly reduce-scatter (lane l ends with column l)
            float s2[32];
#pragma unroll
            for (int jj = 0; jj < 32; ++jj) {
              v[jj] = row_ok ? v[jj] : 0.0f;
              s2[jj] = v[jj] * v[jj];
            }
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) {
              const bool up = (lane & off) != 0;
#pragma unroll
              for (int i = 0; i < off; ++i) {
                const float send1 = up ? v[i] : v[i + off];
                const float keep1 = up ? v[i + off] : v[i];
                v[i] = keep1 + __shfl_xor_sync(0xffffffffu, send1, off);
                const float send2 = up ? s2[i] : s2[i + off];
                const float keep2 = up ? s2[i + off] : s2[i];
                s2[i] = keep2 + __shfl_xor_sync(0xffffffffu, send2, off);
              }
            }
            float* st = stat_s + quarter * 4 * BN + ch * 32 + lane;      // lane l holds column l (half-warp-0 copy)
            st[0] += v[0];
            st[BN] += s2[0];
          }
        }
      };
      uint32_t raw_a[32], raw_b[32];
      if (!chunk_ok(0)) {
        release_acc();
      } else {
        issue(0, raw_a);
#pragma unroll 1
        for (int j = 0; j < kMyChunks; j += 2) {
          if (!chunk_ok(j)) break;
          tc_wait_ld();
          const bool more_b = chunk_ok(j + 1);
          if (more_b) issue(j + 1, raw_b);
          else release_acc();
          process(j, raw_a);
          if (more_b) {
            tc_wait_ld();
            const bool more_a = chunk_ok(j + 2);
            if (more_a) issue(j + 2, raw_a);
            else release_acc();
            process(j + 1, raw_b);
          }
        }
      }
    }
    if (tma_store && lane == 0) bulk_wait_all();   // all bulk stores of this warp have landed
    if (want_stats && cur_nblk >= 0) {
      // ---- final flush (same as above)
      asm volatile("bar.sync 1, 256;" ::: "memory");
      for (int c = epi_tid; c < BN; c += 32 * kEpiWarps) {
        const int col = cur_nblk * BN + c;
        if (col < N) {
          float a1 = 0.f, a2 = 0.f;
#pragma unroll
          for (int qq = 0; qq < 8; ++qq) {
            a1 += stat_s[qq * 2 * BN + c];
            a2 += stat_s[qq * 2 * BN + BN + c];
          }
          atomicAdd(ep.col_sum + col, a1);
          atomicAdd(ep.col_sumsq + col, a2);
        }
      }
    }
    }
  } else if (TF) {
    // =========================== A transform (warps 10..13) ===========================
    // thread = (16-byte chunk c of the 128-byte row, rows rb + 16 i): 8 channels whose scale/shift are
    // loaded once per k-block.  Branch-free and batched (8 LDS.128 in flight, then math, then 8 STS.128);
    // zero-padding taps of an im2col tile keep the 0 the TMA wrote (select, not branch).
    const int tt = threadIdx.x - kTfThread0;
    const int c = tt & 7;
    const int rb = tt >> 3;
    int stage = 0;
    uint32_t phase = 0;
    uint32_t row_off[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int row = rb + 16 * i;
      row_off[i] = (uint32_t)(row * 128 + ((c ^ (row & 7)) << 4));
    }
    for (int tile = tile0; tile < num_tiles; tile += tile_step) {
      const int n_blk = tile / m_blocks;
      const int m_blk = row_block(tile - n_blk * m_blocks);
      int iy0[8], ix0[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        iy0[i] = 0;
        ix0[i] = 0;
      }
      if (g.is_conv) {
        const int pq = g.P * g.Q;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int m = m_blk * BM + rb + 16 * i;
          const int n = m / pq;
          const int rem = m - n * pq;
          const int p = rem / g.Q;
          iy0[i] = (m < M) ? p * g.stride + g.lower_h : -(1 << 20);
          ix0[i] = (rem - p * g.Q) * g.stride + g.lower_w;
        }
      }
      int r = 0, s = 0, slab = 0;
      const unsigned uH = g.is_conv ? (unsigned)g.H : 0x7fffffffu, uW = g.is_conv ? (unsigned)g.W : 0x7fffffffu;
      for (int kb = 0; kb < k_blocks; ++kb) {
        const int c0 = (g.is_conv ? slab : kb0 + kb) * BK + c * 8;
        const float4 sc0 = __ldg(reinterpret_cast<const float4*>(at.scale + c0));
        const float4 sc1 = __ldg(reinterpret_cast<const float4*>(at.scale + c0 + 4));
        const float4 sh0 = __ldg(reinterpret_cast<const float4*>(at.shift + c0));
        const float4 sh1 = __ldg(reinterpret_cast<const float4*>(at.shift + c0 + 4));
        mbar_wait(&full_bar[stage], phase);
        uint8_t* sa = smem + stage * L::kStageBytes;
        uint4 u[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) u[i] = *reinterpret_cast<const uint4*>(sa + row_off[i]);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const bool ok = (unsigned)(iy0[i] + r) < uH && (unsigned)(ix0[i] + s) < uW;
          float2 a0 = make_float2(__uint_as_float(u[i].x << 16), __uint_as_float(u[i].x & 0xffff0000u));
          float2 a1 = make_float2(__uint_as_float(u[i].y << 16), __uint_as_float(u[i].y & 0xffff0000u));
          float2 a2 = make_float2(__uint_as_float(u[i].z << 16), __uint_as_float(u[i].z & 0xffff0000u));
          float2 a3 = make_float2(__uint_as_float(u[i].w << 16), __uint_as_float(u[i].w & 0xffff0000u));
          a0 = __ffma2_rn(a0, make_float2(sc0.x, sc0.y), make_float2(sh0.x, sh0.y));
          a1 = __ffma2_rn(a1, make_float2(sc0.z, sc0.w), make_float2(sh0.z, sh0.w));
          a2 = __ffma2_rn(a2, make_float2(sc1.x, sc1.y), make_float2(sh1.x, sh1.y));
          a3 = __ffma2_rn(a3, make_float2(sc1.z, sc1.w), make_float2(sh1.z, sh1.w));
          uint4 t;
          if (at.relu) {
            t.x = pack_bf16x2_relu(a0.x, a0.y);
            t.y = pack_bf16x2_relu(a1.x, a1.y);
            t.z = pack_bf16x2_relu(a2.x, a2.y);
            t.w = pack_bf16x2_relu(a3.x, a3.y);
          } else {
            t.x = pack_bf16x2(a0.x, a0.y);
            t.y = pack_bf16x2(a1.x, a1.y);
            t.z = pack_bf16x2(a2.x, a2.y);
            t.w = pack_bf16x2(a3.x, a3.y);
          }
          u[i].x = ok ? t.x : 0u;
          u[i].y = ok ? t.y : 0u;
          u[i].z = ok ? t.z : 0u;
          u[i].w = ok ? t.w : 0u;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) *reinterpret_cast<uint4*>(sa + row_off[i]) = u[i];
        fence_proxy_async_smem();        // generic-proxy writes -> visible to the tensor core (async proxy)
        mbar_arrive(&tf_bar[stage]);
        if (g.is_conv && ++slab == g.c_slabs) {
          slab = 0;
          if (++s == g.S) {
            s = 0;
            ++r;
          }
        }
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  }

  tc_fence_before();
  // every thread makes ITS statistics reductions device-visible BEFORE the block barrier: a fence by thread 0
  // alone (after the barrier) does not wait for the other warps' in-flight reductions, and the last CTA then
  // finalised from incomplete sums whenever the CTAs' finish times spread (seen with a second stream running)
  if (ep.fin.scale != nullptr) __threadfence();
  __syncthreads();
  if (ep.fin.scale != nullptr) {
    // BatchNorm finalisation by the last CTA to get here: every CTA's statistics are in global memory
    __shared__ int is_last;
    if (threadIdx.x == 0) is_last = (atomicAdd(ep.fin.counter, 1u) == gridDim.x - 1);
    __syncthreads();
    if (is_last) {
      __threadfence();
      const BnFinal& f = ep.fin;
      for (int c = threadIdx.x; c < N; c += kThreads) {
        const float mean = __ldcg(ep.col_sum + c) * f.inv_count;
        const float var = fmaxf(__ldcg(ep.col_sumsq + c) * f.inv_count - mean * mean, 0.f);
        if (f.running_mean != nullptr) {
          f.running_mean[c] = (1.f - f.momentum) * f.running_mean[c] + f.momentum * mean;
          f.running_var[c] = (1.f - f.momentum) * f.running_var[c] + f.momentum * var * f.unbias;
        }
        const float sc = f.gamma[c] * rsqrtf(var + f.eps);
        f.scale[c] = sc;
        f.shift[c] = f.beta[c] - mean * sc;
      }
    }
  }
  if (PAIR) cluster_sync();        // both CTAs are done with the pair's tensor memory and barriers
  if (warp == 1) {
    tc_fence_after();
    if (PAIR) tc_dealloc_pair(tmem_base, kTmemCols);
    else tc_dealloc(tmem_base, kTmemCols);
  }
  if (CL) cluster_sync();          // no CTA leaves while its peer can still multicast into it / arrive on its barriers
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t,
                                   const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn g_encode_tiled = nullptr;
EncodeIm2colFn g_encode_im2col = nullptr;
std::once_flag g_driver_once;

int load_driver_entry_points() {
  std::call_once(g_driver_once, [] {
    cudaDriverEntryPointQueryResult qres;
    void* fn = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      g_encode_tiled = reinterpret_cast<EncodeTiledFn>(fn);
    fn = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      g_encode_im2col = reinterpret_cast<EncodeIm2colFn>(fn);
  });
  if (g_encode_tiled == nullptr || g_encode_im2col == nullptr) {
    b2_set_error("cuTensorMapEncode* driver entry points unavailable (no CUDA driver / GPU?)");
    return -2;
  }
  return 0;
}

int make_tmap_2d(CUtensorMap* map, const void* base, long rows, long cols, long ld_elems, int box_rows,
                 int box_cols = BK, CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld_elems * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode_tiled(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box,
                              estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    b2_set_error("cuTensorMapEncodeTiled failed (%d): rows=%ld cols=%ld ld=%ld box_rows=%d base=%p", (int)r, rows,
                 cols, ld_elems, box_rows, base);
    return -3;
  }
  return 0;
}

template <int BN, int MODE, bool TF, bool CL = false, bool DEEP = false, bool PAIR = false>
int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& td, const CUtensorMap& tr, int M, int N,
                int K, const ConvGeom& g, const ATransform& at, const EpiParams& ep, cudaStream_t stream) {
  using L = SmemLayout<BN, MODE, DEEP, PAIR>;
  static B2PerDeviceOnce attr_set;
  if (attr_set.needed()) {
    B2_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc_kernel<BN, MODE, TF, CL, DEEP, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       L::kTotal));
    attr_set.mark();
  }
  const int m_blocks = CL ? ((b2_ceil_div(M, BM) + 1) & ~1) : b2_ceil_div(M, BM);
  const int tiles = PAIR ? 2 * ((m_blocks + 1) / 2) * b2_ceil_div(N, BN) : m_blocks * b2_ceil_div(N, BN);   // in CTAs
  int grid = tiles < b2_num_sms() ? tiles : b2_num_sms();
  if (CL || PAIR) grid &= ~1;
  cudaLaunchConfig_t cfg = {};
  const int splits = (MODE == EPI_GENERIC && ep.k_splits > 1) ? ep.k_splits : 1;
  cfg.gridDim = dim3(grid, splits);
  cfg.blockDim = dim3(TF ? kThreadsTf : kThreadsNoTf);
  cfg.dynamicSmemBytes = L::kTotal;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  static const bool pdl = getenv("B2_PDL") != nullptr;   // opt-in: measured no gain (encoder pass 5.02 vs 5.01 ms)
  if (pdl) {        // programmatic dependent launch: start the set-up while the previous kernel drains
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (CL || PAIR) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 2;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  B2_CUDA_CHECK(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN, MODE, TF, CL, DEEP, PAIR>, ta, tb, td, tr, M, N, K, g, at, ep));
  B2_LAUNCH_CHECK("gemm_tc_kernel");
  return 0;
}

int pick_bn(int N) {
  if (N > 128) return 256;
  if (N > 64) return 128;
  if (N > 32) return 64;
  return 32;
}

// Plain GEMMs (Linear layers, their gradients): a small problem in 256-wide tiles leaves most SMs idle -- narrow
// the tile until the grid covers the machine (M = 1024, N = 1024: 32 tiles of 256 -> 128 tiles of 64).
int pick_bn_gemm(int M, int N) {
  int bn = pick_bn(N);
  const int m_blocks = b2_ceil_div(M, BM);
  while (bn > 64 && m_blocks * b2_ceil_div(N, bn) < b2_num_sms()) bn >>= 1;
  return bn;
}

template <int MODE, bool TF>
int dispatch_bn(int bn, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& td, const CUtensorMap& tr,
                int M, int N, int K, const ConvGeom& g, const ATransform& at, const EpiParams& ep, cudaStream_t stream,
                const CUtensorMap* tb_half = nullptr, const CUtensorMap* tb_pair = nullptr) {
  if constexpr (MODE != EPI_GENERIC) {
    if (bn == 256 && tb_half != nullptr)      // two-CTA clusters with the weight tile multicast (see the kernel)
      return launch_gemm<256, MODE, TF, true>(ta, *tb_half, td, tr, M, N, K, g, at, ep, stream);
  }
  if constexpr (MODE != EPI_GENERIC && !TF) {
    // CTA pairs (tcgen05.mma.cta_group::2, M = 256): the wide layers (256+ output channels, K >= 256) -- each CTA stages
    // half of the weight tile, so the shared-memory read and L2 -> SM traffic of the B operand halve
    if (bn == 256 && K >= 256 && tb_pair != nullptr)
      return launch_gemm<256, MODE, TF, false, false, true>(ta, *tb_pair, td, tr, M, N, K, g, at, ep, stream);
  }
  if constexpr ((MODE == EPI_BF16 || MODE == EPI_STATS) && !TF) {
    if (bn == 256 && K >= 512 && getenv("B2_NO_DEEP") == nullptr)   // long K: four operand stages
      return launch_gemm<256, MODE, TF, false, true>(ta, tb, td, tr, M, N, K, g, at, ep, stream);
  }
  switch (bn) {
    case 256: return launch_gemm<256, MODE, TF>(ta, tb, td, tr, M, N, K, g, at, ep, stream);
    case 128: return launch_gemm<128, MODE, TF>(ta, tb, td, tr, M, N, K, g, at, ep, stream);
    case 64: return launch_gemm<64, MODE, TF>(ta, tb, td, tr, M, N, K, g, at, ep, stream);
    default:
      if constexpr (MODE == EPI_GENERIC && !TF)
        return launch_gemm<32, MODE, TF>(ta, tb, td, tr, M, N, K, g, at, ep, stream);
      b2_set_error("gemm_tc: this epilogue needs at least 64 output columns (got %d)", N);
      return -1;
  }
}

int dispatch(int bn, const CUtensorMap& ta, const CUtensorMap& tb, int M, int N, int K, const ConvGeom& g,
             const ATransform& at, EpiParams ep, cudaStream_t stream, const CUtensorMap* tb_half = nullptr,
             const CUtensorMap* tb_pair = nullptr) {
  // bf16 outputs whose rows are 16-byte multiples leave through TMA bulk stores
  CUtensorMap td = ta, tr = ta;
  ep.tma_store = 0;
  if (ep.store && ep.out_bf16 && (ep.ldd % 8) == 0 && ((uintptr_t)ep.D & 15) == 0) {
    if (int r = make_tmap_2d(&td, ep.D, M, N, ep.ldd, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B)) return r;
    ep.tma_store = 1;
  }
  if (ep.res != nullptr) {   // shortcut tiles are fetched by TMA (same 32x32 SWIZZLE_64B boxes as the store)
    if (int r = make_tmap_2d(&tr, ep.res, M, N, ep.ldres, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B)) return r;
  }
  // pick the compile-time specialisation: the lean conv epilogues need whole 32-column chunks, a bf16 output
  // that can leave by TMA, and no bias / ReLU on the raw output; everything else takes the generic one
  const bool tf = at.scale != nullptr;
  const bool lean = bn >= 64 && (N % 32) == 0 && ep.bias == nullptr && ep.bias2 == nullptr && ep.out_bf16;
  const bool post = ep.o_scale != nullptr || ep.res != nullptr;
  int mode = EPI_GENERIC;
  if (lean && !ep.store && !post) mode = EPI_STATS;
  else if (lean && ep.store && ep.tma_store && post) mode = EPI_POST;
  else if (lean && ep.store && ep.tma_store && !ep.relu) mode = EPI_BF16;
  if (mode == EPI_GENERIC && (tf || post || !ep.store)) {
    b2_set_error("gemm_tc: BatchNorm folding needs a 16 B aligned bf16 output with N %% 32 == 0 and N >= 64 (N=%d)", N);
    return -1;
  }
  switch (mode) {
    case EPI_STATS:
      return tf ? dispatch_bn<EPI_STATS, true>(bn, ta, tb, td, tr, M, N, K, g, at, ep, stream, tb_half)
                : dispatch_bn<EPI_STATS, false>(bn, ta, tb, td, tr, M, N, K, g, at, ep, stream, tb_half, tb_pair);
    case EPI_POST:
      return tf ? dispatch_bn<EPI_POST, true>(bn, ta, tb, td, tr, M, N, K, g, at, ep, stream, tb_half)
                : dispatch_bn<EPI_POST, false>(bn, ta, tb, td, tr, M, N, K, g, at, ep, stream, tb_half, tb_pair);
    case EPI_BF16:
      return tf ? dispatch_bn<EPI_BF16, true>(bn, ta, tb, td, tr, M, N, K, g, at, ep, stream, tb_half)
                : dispatch_bn<EPI_BF16, false>(bn, ta, tb, td, tr, M, N, K, g, at, ep, stream, tb_half, tb_pair);
    default:
      return dispatch_bn<EPI_GENERIC, false>(bn, ta, tb, td, tr, M, N, K, g, at, ep, stream);
  }
}

EpiParams plain_epi(void* D, long ldd, const float* bias, const float* bias2, int out_bf16, int relu, float* col_sum,
                    float* col_sumsq) {
  EpiParams ep = {};
  ep.D = D;
  ep.ldd = ldd;
  ep.bias = bias;
  ep.bias2 = bias2;
  ep.col_sum = col_sum;
  ep.col_sumsq = col_sumsq;
  ep.out_bf16 = out_bf16;
  ep.relu = relu;
  ep.store = 1;
  return ep;
}

int conv_common(const void* x, int Nimg, int H, int W, int C, const void* w, int Cout, int R, int S, int stride,
                int pad, const ATransform& at, EpiParams ep, cudaStream_t stream, const char* who) {
  B2_ARG_CHECK(x && w && Nimg > 0 && H > 0 && W > 0 && Cout > 0, "%s: null pointer or empty shape", who);
  B2_ARG_CHECK(C % 64 == 0, "%s: C must be a multiple of 64 (got %d)", who, C);
  B2_ARG_CHECK(R >= 1 && S >= 1 && R <= 7 && S <= 7 && stride >= 1 && stride <= 8 && pad >= 0 && pad <= 3,
               "%s: unsupported filter geometry R=%d S=%d stride=%d pad=%d", who, R, S, stride, pad);
  B2_ARG_CHECK((ep.col_sum == nullptr) == (ep.col_sumsq == nullptr), "%s: col_sum and col_sumsq go together", who);
  const int P = (H + 2 * pad - R) / stride + 1;
  const int Q = (W + 2 * pad - S) / stride + 1;
  B2_ARG_CHECK(P > 0 && Q > 0, "%s: empty output", who);
  if (int r = load_driver_entry_points()) return r;
  const long Ml = (long)Nimg * P * Q;
  B2_ARG_CHECK(Ml < (1L << 31), "%s: too many output pixels", who);
  const int M = (int)Ml;
  const int K = R * S * C;
  const int bn = pick_bn(Cout);
  ep.ldd = Cout;
  if (ep.res != nullptr) ep.ldres = Cout;
  CUtensorMap ta, tb, tbh;
  if (int r = make_tmap_2d(&tb, w, Cout, K, K, bn)) return r;
  // weight-multicast clusters pay off where the B tile dominates the operand traffic and the grid fills the machine
  const CUtensorMap* tb_half = nullptr;
  if (bn == 256 && K >= 256 && Cout % 256 == 0 && (long)b2_ceil_div(M, BM) * (Cout / 256) >= b2_num_sms() &&
      getenv("B2_CLUSTER") != nullptr) {   // opt-in: measured no gain on B200 (unicast TMA from neighbouring SMs is already de-duplicated in L2)
    if (int r = make_tmap_2d(&tbh, w, Cout, K, K, bn / 2)) return r;
    tb_half = &tbh;
  }
  // CTA pairs (cta_group::2): every wide conv without an A transform; B2_PAIR=0 falls back to one CTA per tile (A/B runs)
  static const bool pair_on = getenv("B2_PAIR") == nullptr || getenv("B2_PAIR")[0] != '0';
  CUtensorMap tbp;
  const CUtensorMap* tb_pair = nullptr;
  if (pair_on && tb_half == nullptr && bn == 256 && K >= 256 && Cout % 256 == 0 && at.scale == nullptr) {
    if (int r = make_tmap_2d(&tbp, w, Cout, K, K, bn / 2)) return r;
    tb_pair = &tbp;
  }
  if (R == 1 && S == 1 && stride == 1 && pad == 0) {
    if (int r = make_tmap_2d(&ta, x, M, C, C, BM)) return r;
    ConvGeom g = {};
    return dispatch(bn, ta, tb, M, Cout, K, g, at, ep, stream, tb_half, tb_pair);
  }
  // IM2COL-mode map over the NHWC activation: dims {C, W, H, N}; the bounding box of filter-window
  // base positions is [-pad, dim-1 + (pad - (R-1))] and is walked with the convolution stride.
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)Nimg};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  int lower[2] = {-pad, -pad};
  int upper[2] = {pad - (S - 1), pad - (R - 1)};
  cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
  CUresult cr = g_encode_im2col(&ta, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), dims, strides, lower,
                                upper, (cuuint32_t)BK, (cuuint32_t)BM, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) {
    b2_set_error("cuTensorMapEncodeIm2col failed (%d): N=%d H=%d W=%d C=%d R=%d S=%d stride=%d pad=%d", (int)cr, Nimg,
                 H, W, C, R, S, stride, pad);
    return -3;
  }
  // Driver quirk (CUDA <= 13.1): im2col maps over tensors smaller than 128 KiB need bit 21 of the
  // second descriptor word cleared, otherwise loads fault.
  {
    int drv = 0;
    cudaDriverGetVersion(&drv);
    if (drv <= 13010 && (long)Nimg * H * W * C * 2 < 131072) reinterpret_cast<uint64_t*>(&ta)[1] &= ~(1ull << 21);
  }
  ConvGeom g = {1, P, Q, S, C / 64, stride, -pad, -pad, H, W};
  return dispatch(bn, ta, tb, M, Cout, K, g, at, ep, stream, tb_half, tb_pair);
}

// ------------------------------------------------------------------ Gram-matrix BatchNorm statistics
// Train-mode BN3 of a bottleneck needs the statistics of the conv3 output y = t W^T (t = relu(bn2(conv2 raw)),
// [M, K]; W [N, K], N = 4K) before anything can be normalised.  A statistics-only GEMM pass drains M x N
// accumulators from TMEM (64 B/clk/SM) and is epilogue bound.  Both moments follow from K-sized reductions:
//     sum_m y[m,n]   = < w_n , s >          s = sum_m t[m,:]
//     sum_m y[m,n]^2 = w_n^T G w_n          G = sum_m t[m,:]^T t[m,:]   (K x K Gram matrix)
// so this kernel streams the activation ONCE (HBM bound), rewrites each tile in shared memory exactly as the
// conv3 A transform does, and feeds the SAME tile to the tensor core as both operands in MN-major form
// (reduction over the 128 rows of the tile): G accumulates in TMEM over all of a CTA's tiles and is drained
// once at the end (vector reductions into a global fp32 workspace); s comes from one extra N = 16 MMA against
// a ones matrix.  gram_finalize_kernel evaluates the two forms per output channel in fp64 and finalises the
// BatchNorm.  Exact for the unrounded fp32 conv output (more accurate than summing bf16-rounded outputs).
//
// A "super tile" is two 16 KB SWIZZLE_128B panels of 128 rows x 64 channels, M/N = 128 for the MMA:
//   K = 128: the two 64-channel halves of one 128-row tile;  K = 64: two consecutive 128-row tiles side by
//   side (the Gram of [t1 | t2] holds t1^T t1 and t2^T t2 on its diagonal blocks, summed at drain time).
constexpr int kGramCopies = 8;           // partial global accumulators: shortens the same-address atomic chains 8x
constexpr int kGramThreads = 64 + 256;   // TMA warp, MMA warp, 8 transform warps (the first 4 also drain)
constexpr int kPanelBytes = BM * BK * 2; // 16 KB: 128 rows x 64 channels, SWIZZLE_128B

template <int KC>
struct GramLayout {
  static constexpr int kPanels = KC == 256 ? 4 : 2;              // panels per super tile
  static constexpr int kStages = KC == 256 ? 3 : 5;
  static constexpr int kStageBytes = kPanels * kPanelBytes;      // 32 / 64 KB
  static constexpr int kOnesOffset = kStages * kStageBytes;      // 1 KB of bf16 ones
  static constexpr int kBarOffset = kOnesOffset + 1024;
  static constexpr int kSmem = kBarOffset + (3 * kStages + 1) * 8 + 16 + 1024;
  // TMEM columns: K <= 128: G [0,128), s [128,144).  K = 256: G rows 0..127 x all 256 columns [0,256),
  // G rows 128..255 x columns 128..255 [256,384) (the lower-left block is the transpose of the upper-right one),
  // s rows 0..127 [384,400), s rows 128..255 [400,416)
  static constexpr uint32_t kTmemCols = KC == 256 ? 512 : 256;
  static constexpr uint32_t kSumCol0 = KC == 256 ? 384 : 128;
};

// MN-major SWIZZLE_128B descriptor: 64 contiguous MN elements per 128 B row (one row per K index), 8-row
// groups SBO = 1024 B apart, 64-element MN blocks LBO bytes apart (the panel stride).
__device__ __forceinline__ uint64_t make_sw128_mn_desc(uint32_t smem_addr, uint32_t lbo) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

template <int KC>
__global__ void __launch_bounds__(kGramThreads, 1)
gram_stats_kernel(const __grid_constant__ CUtensorMap tmap_a, int M, ATransform at, float* __restrict__ ws) {
  static_assert(KC == 64 || KC == 128 || KC == 256, "gram_stats_kernel: 64, 128 or 256 input channels");
  using G = GramLayout<KC>;
  constexpr int kStages = G::kStages;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + G::kBarOffset);
  uint64_t* tf_bar = full_bar + kStages;
  uint64_t* empty_bar = tf_bar + kStages;
  uint64_t* done_bar = empty_bar + kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done_bar + 1);
  constexpr int kRowsPerSuper = KC == 64 ? 2 * BM : BM;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_super = (M + kRowsPerSuper - 1) / kRowsPerSuper;

  for (int i = threadIdx.x; i < 256; i += kGramThreads)
    reinterpret_cast<uint32_t*>(smem + G::kOnesOffset)[i] = 0x3F803F80u;   // bf16 1.0 pairs
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&tf_bar[i], 256);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(done_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tc_alloc(tmem_slot, G::kTmemCols);
    tc_relinquish();
  }
  fence_proxy_async_smem();        // the ones matrix is read by the tensor core (async proxy)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int st = blockIdx.x; st < num_super; st += gridDim.x) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * G::kStageBytes;
        mbar_expect_tx(&full_bar[stage], G::kStageBytes);
        if (KC == 64) {
          tma_load_2d(sa, &tmap_a, &full_bar[stage], 0, st * 2 * BM);
          tma_load_2d(sa + kPanelBytes, &tmap_a, &full_bar[stage], 0, st * 2 * BM + BM);   // fully OOB rows -> zeros
        } else {
#pragma unroll
          for (int pp = 0; pp < G::kPanels; ++pp)
            tma_load_2d(sa + pp * kPanelBytes, &tmap_a, &full_bar[stage], pp * BK, st * BM);
        }
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // operands MN-major (bits 15, 16): D += P^T P ; column sums: D[128 x 16] += P^T ones (B K-major)
    constexpr uint32_t idesc_g = make_idesc(128, KC == 256 ? 256 : 128) | (1u << 15) | (1u << 16);
    constexpr uint32_t idesc_g2 = make_idesc(128, 128) | (1u << 15) | (1u << 16);
    constexpr uint32_t idesc_s = make_idesc(128, 16) | (1u << 15);
    const uint64_t ones_desc = make_nosw_desc(smem_u32(smem + G::kOnesOffset), 128, 256);
    int stage = 0;
    uint32_t phase = 0;
    bool first = true;
    for (int st = blockIdx.x; st < num_super; st += gridDim.x) {
      mbar_wait(&tf_bar[stage], phase);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t sa = smem_u32(smem + stage * G::kStageBytes);
#pragma unroll
        for (int ks = 0; ks < BM / UMMA_K; ++ks) {          // 16 tile rows (two 8-row groups) per MMA
          const uint32_t acc = !(first && ks == 0);
          const uint64_t d0 = make_sw128_mn_desc(sa + ks * 2048, kPanelBytes);
          tc_mma_bf16(tmem_base, d0, d0, idesc_g, acc);                       // features 0..127 x all
          tc_mma_bf16(tmem_base + G::kSumCol0, d0, ones_desc, idesc_s, acc);
          if (KC == 256) {
            const uint64_t d2 = make_sw128_mn_desc(sa + 2 * kPanelBytes + ks * 2048, kPanelBytes);
            tc_mma_bf16(tmem_base + 256, d2, d2, idesc_g2, acc);              // features 128..255 x 128..255
            tc_mma_bf16(tmem_base + G::kSumCol0 + 16, d2, ones_desc, idesc_s, acc);
          }
        }
        tc_commit(&empty_bar[stage]);
      }
      first = false;
      __syncwarp();
      if (++stage == kStages) {
        stage = 0;
        phase ^= 1;
      }
    }
    if (lane == 0) tc_commit(done_bar);
    __syncwarp();
  } else {
    // =========================== A transform (warps 2..9), then the drain (warps 2..5) ===========================
    const int tt = (threadIdx.x - 64) & 127;
    const int grp = (threadIdx.x - 64) >> 7;        // group of 4 warps: panels grp, grp + 2 (K = 256) or panel grp
    const int c = tt & 7;
    const int rb = tt >> 3;
    uint32_t row_off[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int row = rb + 16 * i;
      row_off[i] = (uint32_t)(row * 128 + ((c ^ (row & 7)) << 4));
    }
    constexpr int kMine = KC == 256 ? 2 : 1;        // panels per thread group; a thread's channels are fixed per panel
    float4 sc0[kMine], sc1[kMine], sh0[kMine], sh1[kMine];
#pragma unroll
    for (int q = 0; q < kMine; ++q) {
      const int c0 = (KC == 64 ? 0 : (grp + 2 * q) * BK) + c * 8;
      sc0[q] = __ldg(reinterpret_cast<const float4*>(at.scale + c0));
      sc1[q] = __ldg(reinterpret_cast<const float4*>(at.scale + c0 + 4));
      sh0[q] = __ldg(reinterpret_cast<const float4*>(at.shift + c0));
      sh1[q] = __ldg(reinterpret_cast<const float4*>(at.shift + c0 + 4));
    }
    int stage = 0;
    uint32_t phase = 0;
    for (int st = blockIdx.x; st < num_super; st += gridDim.x) {
      mbar_wait(&full_bar[stage], phase);
#pragma unroll
      for (int q = 0; q < kMine; ++q) {
        const int p = grp + 2 * q;
        uint8_t* sa = smem + stage * G::kStageBytes + p * kPanelBytes;
        const int row0 = (KC == 64 ? st * 2 * BM + p * BM : st * BM) + rb;
        uint4 u[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) u[i] = *reinterpret_cast<const uint4*>(sa + row_off[i]);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const bool ok = row0 + 16 * i < M;           // rows past the end stay zero
          float2 a0 = make_float2(__uint_as_float(u[i].x << 16), __uint_as_float(u[i].x & 0xffff0000u));
          float2 a1 = make_float2(__uint_as_float(u[i].y << 16), __uint_as_float(u[i].y & 0xffff0000u));
          float2 a2 = make_float2(__uint_as_float(u[i].z << 16), __uint_as_float(u[i].z & 0xffff0000u));
          float2 a3 = make_float2(__uint_as_float(u[i].w << 16), __uint_as_float(u[i].w & 0xffff0000u));
          a0 = __ffma2_rn(a0, make_float2(sc0[q].x, sc0[q].y), make_float2(sh0[q].x, sh0[q].y));
          a1 = __ffma2_rn(a1, make_float2(sc0[q].z, sc0[q].w), make_float2(sh0[q].z, sh0[q].w));
          a2 = __ffma2_rn(a2, make_float2(sc1[q].x, sc1[q].y), make_float2(sh1[q].x, sh1[q].y));
          a3 = __ffma2_rn(a3, make_float2(sc1[q].z, sc1[q].w), make_float2(sh1[q].z, sh1[q].w));
          uint4 t;
          if (at.relu) {
            t.x = pack_bf16x2_relu(a0.x, a0.y);
            t.y = pack_bf16x2_relu(a1.x, a1.y);
            t.z = pack_bf16x2_relu(a2.x, a2.y);
            t.w = pack_bf16x2_relu(a3.x, a3.y);
          } else {
            t.x = pack_bf16x2(a0.x, a0.y);
            t.y = pack_bf16x2(a1.x, a1.y);
            t.z = pack_bf16x2(a2.x, a2.y);
            t.w = pack_bf16x2(a3.x, a3.y);
          }
          u[i].x = ok ? t.x : 0u;
          u[i].y = ok ? t.y : 0u;
          u[i].z = ok ? t.z : 0u;
          u[i].w = ok ? t.w : 0u;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) *reinterpret_cast<uint4*>(sa + row_off[i]) = u[i];
      }
      fence_proxy_async_smem();
      mbar_arrive(&tf_bar[stage]);
      if (++stage == kStages) {
        stage = 0;
        phase ^= 1;
      }
    }
    // ---- drain (warps 2..5): lane = feature row f of the accumulator; vector reductions into this CTA's
    // partial workspace [KC*KC Gram | KC sums]
    if (warp < 6) {
      mbar_wait(done_bar, 0);
      tc_fence_after();
      const int quarter = warp & 3;
      const int f = quarter * 32 + lane;
      float* gram = ws + (long)(blockIdx.x % kGramCopies) * (KC * KC + KC);
      float* sum_a = gram + KC * KC;
      constexpr int kChunks = KC == 256 ? 12 : 4;
#pragma unroll 1
      for (int ch = 0; ch < kChunks; ++ch) {
        if (KC == 64 && (ch >> 1) != (quarter >> 1)) continue;      // only the two diagonal 64 x 64 blocks
        uint32_t raw[32];
        tc_ld32(tmem_base + (uint32_t)(ch * 32) + ((uint32_t)(quarter * 32) << 16), raw);
        tc_wait_ld();
        float* dst;
        if (KC == 64) dst = gram + (long)(f & 63) * 64 + (ch & 1) * 32;
        else if (KC == 128) dst = gram + (long)f * 128 + ch * 32;
        else dst = ch < 8 ? gram + (long)f * 256 + ch * 32 : gram + (long)(128 + f) * 256 + 128 + (ch - 8) * 32;
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j), "f"(__uint_as_float(raw[j])),
                       "f"(__uint_as_float(raw[j + 1])), "f"(__uint_as_float(raw[j + 2])), "f"(__uint_as_float(raw[j + 3]))
                       : "memory");
      }
      {
        uint32_t raw[32];     // every column of an N = 16 block holds s; K = 256: columns 0 and 16 of this load
        tc_ld32(tmem_base + G::kSumCol0 + ((uint32_t)(quarter * 32) << 16), raw);
        tc_wait_ld();
        atomicAdd(sum_a + (KC == 64 ? (f & 63) : f), __uint_as_float(raw[0]));
        if (KC == 256) atomicAdd(sum_a + 128 + f, __uint_as_float(raw[16]));
      }
      tc_fence_before();
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tc_dealloc(tmem_base, G::kTmemCols);
  }
}

// Finalisation, two small kernels:
//   gram_reduce_kernel   sums the kGramCopies partial workspaces, mirrors the block the K = 256 kernel leaves out
//                        and centres: Cov = G/M - m m^T, m = s/M
//   gram_finalize_kernel one warp per pair of output channels: mean_n = <w_n, m>, var_n = w_n^T Cov w_n in fp32
//                        (the centred form keeps the cancellation out of the long sums: 1e-5 relative on var
//                        against an fp64 evaluation at the layer shapes), then the BatchNorm finalisation.
__global__ void __launch_bounds__(256)
gram_reduce_kernel(const float* __restrict__ ws, float* __restrict__ cov, int KC, float inv) {
  const int copy_stride = KC * KC + KC;
  const int i4 = blockIdx.x * blockDim.x + threadIdx.x;          // one float4 of the K x K matrix per thread
  if (i4 >= KC * KC / 4) return;
  const int k = (i4 * 4) / KC, l = (i4 * 4) - k * KC;
  const bool mirror = KC == 256 && k >= 128 && l < 128;           // G[k][l] = G[l][k]
  float4 t = make_float4(0.f, 0.f, 0.f, 0.f), ml = t;
  float mk = 0.f;
#pragma unroll
  for (int c = 0; c < kGramCopies; ++c) {
    const float* base = ws + c * copy_stride;
    float4 g;
    if (mirror) {
      g.x = __ldg(base + (l + 0) * KC + k);
      g.y = __ldg(base + (l + 1) * KC + k);
      g.z = __ldg(base + (l + 2) * KC + k);
      g.w = __ldg(base + (l + 3) * KC + k);
    } else {
      g = __ldg(reinterpret_cast<const float4*>(base) + i4);
    }
    const float4 sl = __ldg(reinterpret_cast<const float4*>(base + KC * KC + l));
    mk += __ldg(base + KC * KC + k);
    t.x += g.x; t.y += g.y; t.z += g.z; t.w += g.w;
    ml.x += sl.x; ml.y += sl.y; ml.z += sl.z; ml.w += sl.w;
  }
  mk *= inv;
  t.x = fmaf(t.x, inv, -mk * (ml.x * inv));
  t.y = fmaf(t.y, inv, -mk * (ml.y * inv));
  t.z = fmaf(t.z, inv, -mk * (ml.z * inv));
  t.w = fmaf(t.w, inv, -mk * (ml.w * inv));
  reinterpret_cast<float4*>(cov)[i4] = t;
  if (l == 0) cov[KC * KC + k] = mk;                              // the mean vector follows the matrix
}

constexpr int kGramFinWarps = 8;
constexpr int kGramFinPerWarp = 2;       // output channels per warp
constexpr int kGramFinRows = 64;         // Cov rows per CTA (grid.y = K / 64 row chunks: partial quadratic forms)
// grid = (ceil(N / 16), K / 64).  CTA (nb, kc) stages Cov rows [64 kc, 64 kc + 64) and adds
//   q_n += sum_{k in chunk} sum_l w_k Cov[k][l] w_l      (the quadratic form is linear in the row chunk)
// into acc[n]; the last chunk-CTA of a channel block (counter) turns (mean, var) into the BatchNorm coefficients.
// acc [N] and counters [gridDim.x] are zeroed by the caller.
__global__ void __launch_bounds__(32 * kGramFinWarps)
gram_finalize_kernel(const float* __restrict__ cov, const bf16* __restrict__ W, int N, int KC, float* __restrict__ acc,
                     unsigned int* __restrict__ counters, float* __restrict__ col_sum, float* __restrict__ col_sumsq,
                     BnFinal f, float count) {
  extern __shared__ float gsm[];             // Cov rows [64][KC], mean [KC], weight rows [warps][2][KC]
  float* cov_s = gsm;
  float* m_s = gsm + kGramFinRows * KC;
  float* w_s = m_s + KC;
  __shared__ int is_last;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = (blockIdx.x * kGramFinWarps + warp) * kGramFinPerWarp;
  const int k0 = blockIdx.y * kGramFinRows;
  const int rows = min(kGramFinRows, KC - k0);
  // staging: batches of 8 independent float4 loads per thread
  const float4* src = reinterpret_cast<const float4*>(cov + (long)k0 * KC);
  const int n4 = rows * KC / 4;
  for (int i0 = 0; i0 < n4; i0 += 8 * 32 * kGramFinWarps) {
    float4 v[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int i = i0 + q * 32 * kGramFinWarps + threadIdx.x;
      v[q] = i < n4 ? __ldg(src + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int i = i0 + q * 32 * kGramFinWarps + threadIdx.x;
      if (i < n4) reinterpret_cast<float4*>(cov_s)[i] = v[q];
    }
  }
  float* w0 = w_s + (warp * kGramFinPerWarp) * KC;
  float* w1 = w0 + KC;
  for (int l = lane; l < KC; l += 32) {
    w0[l] = (n0 < N) ? __bfloat162float(W[(long)n0 * KC + l]) : 0.f;
    w1[l] = (n0 + 1 < N) ? __bfloat162float(W[(long)(n0 + 1) * KC + l]) : 0.f;
  }
  for (int i = threadIdx.x; i < KC; i += blockDim.x) m_s[i] = cov[KC * KC + i];
  __syncthreads();
  const int per = KC / 32;                   // 2, 4 or 8 columns of Cov per lane
  float u0[8], u1[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) u0[j] = u1[j] = 0.f;
#pragma unroll 4
  for (int k = 0; k < rows; ++k) {
    const float a0 = w0[k0 + k], a1 = w1[k0 + k];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (j < per) {
        const float cv = cov_s[k * KC + lane + 32 * j];
        u0[j] = fmaf(a0, cv, u0[j]);
        u1[j] = fmaf(a1, cv, u1[j]);
      }
    }
  }
  float q0 = 0.f, q1 = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    if (j < per) {
      const int l = lane + 32 * j;
      q0 = fmaf(w0[l], u0[j], q0);
      q1 = fmaf(w1[l], u1[j], q1);
    }
  }
  q0 = warp_sum(q0);
  q1 = warp_sum(q1);
  if (lane < kGramFinPerWarp && n0 + lane < N) atomicAdd(acc + n0 + lane, lane == 0 ? q0 : q1);
  // ---- the last row-chunk CTA of this channel block finalises its channels
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = (atomicAdd(counters + blockIdx.x, 1u) == gridDim.y - 1);
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  float m0 = 0.f, m1 = 0.f;
  for (int l = lane; l < KC; l += 32) {
    m0 = fmaf(w0[l], m_s[l], m0);
    m1 = fmaf(w1[l], m_s[l], m1);
  }
  m0 = warp_sum(m0);
  m1 = warp_sum(m1);
  if (lane < kGramFinPerWarp && n0 + lane < N) {
    const int n = n0 + lane;
    const float mean = lane == 0 ? m0 : m1;
    const float var = fmaxf(__ldcg(acc + n), 0.f);
    if (col_sum != nullptr) {
      col_sum[n] = mean * count;
      col_sumsq[n] = (var + mean * mean) * count;
    }
    if (f.running_mean != nullptr) {
      f.running_mean[n] = (1.f - f.momentum) * f.running_mean[n] + f.momentum * mean;
      f.running_var[n] = (1.f - f.momentum) * f.running_var[n] + f.momentum * var * f.unbias;
    }
    const float sc = f.gamma[n] * rsqrtf(var + f.eps);
    f.scale[n] = sc;
    f.shift[n] = f.beta[n] - mean * sc;
  }
}

__global__ void bn_finalize_kernel(const float* sum, const float* sumsq, const float* gamma, const float* beta,
                                   float* running_mean, float* running_var, float inv_count, float unbias, float eps,
                                   float momentum, int train, float* scale, float* shift, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float mean, var;
  if (train) {
    mean = sum[c] * inv_count;
    var = fmaxf(sumsq[c] * inv_count - mean * mean, 0.f);
    if (running_mean != nullptr) {
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * var * unbias;
    }
  } else {
    mean = running_mean[c];
    var = running_var[c];
  }
  const float sc = gamma[c] * rsqrtf(var + eps);
  scale[c] = sc;
  shift[c] = beta[c] - mean * sc;
}

}  // namespace

// D[M,N] = A[M,K] B[N,K]^T (+bias) ; see include/b200lrcn.h
B2_API int b2_gemm_bf16_tn(const void* A, long lda, const void* B, long ldb, void* D, long ldd, int M, int N, int K,
                           const float* bias, const float* bias2, int out_bf16, int relu, float* col_sum,
                           float* col_sumsq, void* stream) {
  B2_ARG_CHECK(A && B && D && M > 0 && N > 0 && K > 0, "b2_gemm_bf16_tn: null pointer or empty shape");
  B2_ARG_CHECK((lda % 8) == 0 && (ldb % 8) == 0, "b2_gemm_bf16_tn: lda/ldb must be multiples of 8 elements (16 B)");
  B2_ARG_CHECK(((uintptr_t)A & 15) == 0 && ((uintptr_t)B & 15) == 0, "b2_gemm_bf16_tn: A/B must be 16 B aligned");
  B2_ARG_CHECK((col_sum == nullptr) == (col_sumsq == nullptr), "b2_gemm_bf16_tn: col_sum and col_sumsq go together");
  if (int r = load_driver_entry_points()) return r;
  const int bn = pick_bn_gemm(M, N);
  CUtensorMap ta, tb;
  if (int r = make_tmap_2d(&ta, A, M, K, lda, BM)) return r;
  if (int r = make_tmap_2d(&tb, B, N, K, ldb, bn)) return r;
  ConvGeom g = {};
  ATransform at = {};
  EpiParams ep = plain_epi(D, ldd, bias, bias2, out_bf16, relu, col_sum, col_sumsq);
  // skinny output, long reduction (the hoisted LSTM gate GEMM X[B*T,16384] W_ih^T[16384,4H]): split K over the idle SMs
  const int tiles = b2_ceil_div(M, BM) * b2_ceil_div(N, bn);
  const int k_blocks = b2_ceil_div(K, BK);
  // (K >= 8192 only: the adapt / head Linears keep one CTA per tile, so their forward stays bit-reproducible run to run --
  //  the CUDA-graph replay tests compare graph and eager outputs bit for bit)
  if (!out_bf16 && !relu && col_sum == nullptr && tiles * 2 <= b2_num_sms() && k_blocks >= 128 && getenv("B2_NO_SPLITK") == nullptr) {
    int splits = b2_num_sms() / tiles;
    if (splits > k_blocks / 8) splits = k_blocks / 8;
    const int per = b2_ceil_div(k_blocks, splits);
    splits = b2_ceil_div(k_blocks, per);                 // every split owns at least one k-block
    if (splits > 1) {
      ep.k_splits = splits;
      B2_CUDA_CHECK(cudaMemset2DAsync(D, (size_t)ldd * 4, 0, (size_t)N * 4, (size_t)M, (cudaStream_t)stream));
    }
  }
  return dispatch(bn, ta, tb, M, N, K, g, at, ep, (cudaStream_t)stream);
}

// D[M,N] (bf16) = relu?(A[M,K] * a_scale[k] + a_shift[k]) B[N,K]^T with the pre-activation BatchNorm of a DenseNet layer
// applied to the A tile in shared memory (A row-strided: a channel slice of a concatenated block buffer) ; see the header
B2_API int b2_gemm_bn_bf16_tn(const void* A, long lda, const void* B, long ldb, void* D, long ldd, int M, int N, int K,
                              const float* a_scale, const float* a_shift, int a_relu, float* col_sum, float* col_sumsq,
                              void* stream) {
  const char* who = "b2_gemm_bn_bf16_tn";
  B2_ARG_CHECK(A && B && D && a_scale && a_shift && M > 0 && N > 0 && K > 0, "%s: null pointer or empty shape", who);
  B2_ARG_CHECK((lda % 8) == 0 && (ldb % 8) == 0 && (ldd % 8) == 0, "%s: row strides must be multiples of 8 elements", who);
  B2_ARG_CHECK(((uintptr_t)A & 15) == 0 && ((uintptr_t)B & 15) == 0 && ((uintptr_t)D & 15) == 0, "%s: 16 B alignment", who);
  B2_ARG_CHECK(N % 32 == 0 && N >= 64, "%s: N must be a multiple of 32, at least 64 (got %d)", who, N);
  B2_ARG_CHECK((col_sum == nullptr) == (col_sumsq == nullptr), "%s: col_sum and col_sumsq go together", who);
  if (int r = load_driver_entry_points()) return r;
  const int bn = pick_bn(N);
  CUtensorMap ta, tb;
  if (int r = make_tmap_2d(&ta, A, M, K, lda, BM)) return r;
  if (int r = make_tmap_2d(&tb, B, N, K, ldb, bn)) return r;
  ConvGeom g = {};
  ATransform at = {a_scale, a_shift, a_relu};      // scale / shift must be readable up to the next multiple of 64 channels
  return dispatch(bn, ta, tb, M, N, K, g, at, plain_epi(D, ldd, nullptr, nullptr, 1, 0, col_sum, col_sumsq),
                  (cudaStream_t)stream);
}

// y[N,P,Q,Cout] = conv(x[N,H,W,C], w[Cout,R,S,C]) ; NHWC bf16, C % 64 == 0 ; see include/b200lrcn.h
B2_API int b2_conv2d_nhwc_bf16(const void* x, int Nimg, int H, int W, int C, const void* w, int Cout, int R, int S,
                               int stride, int pad, void* y, const float* bias, int out_bf16, int relu,
                               float* col_sum, float* col_sumsq, void* stream) {
  B2_ARG_CHECK(y != nullptr, "b2_conv2d_nhwc_bf16: null output");
  ATransform at = {};
  return conv_common(x, Nimg, H, W, C, w, Cout, R, S, stride, pad, at,
                     plain_epi(y, Cout, bias, nullptr, out_bf16, relu, col_sum, col_sumsq), (cudaStream_t)stream,
                     "b2_conv2d_nhwc_bf16");
}

// Convolution with the BatchNorms of a ResNet block folded in ; see include/b200lrcn.h
B2_API int b2_conv2d_bn_nhwc_bf16(const void* x, int Nimg, int H, int W, int C, const void* w, int Cout, int R,
                                  int S, int stride, int pad, void* y, const float* a_scale, const float* a_shift,
                                  int a_relu, const float* o_scale, const float* o_shift, const void* res,
                                  const float* r_scale, const float* r_shift, int relu, float* col_sum,
                                  float* col_sumsq, const float* fin_gamma, const float* fin_beta,
                                  float* fin_running_mean, float* fin_running_var, float* fin_scale,
                                  float* fin_shift, unsigned int* fin_counter, float eps, float momentum,
                                  void* stream) {
  const char* who = "b2_conv2d_bn_nhwc_bf16";
  B2_ARG_CHECK((a_scale == nullptr) == (a_shift == nullptr), "%s: a_scale and a_shift go together", who);
  B2_ARG_CHECK((o_scale == nullptr) == (o_shift == nullptr), "%s: o_scale and o_shift go together", who);
  B2_ARG_CHECK((r_scale == nullptr) == (r_shift == nullptr), "%s: r_scale and r_shift go together", who);
  B2_ARG_CHECK(r_scale == nullptr || res != nullptr, "%s: shortcut BatchNorm without a shortcut tensor", who);
  B2_ARG_CHECK(y != nullptr || (o_scale == nullptr && res == nullptr && col_sum != nullptr),
               "%s: a statistics-only pass (y = NULL) takes col_sum/col_sumsq and no output BatchNorm / shortcut", who);
  B2_ARG_CHECK((res == nullptr && o_scale == nullptr) || Cout % 32 == 0,
               "%s: output BatchNorm / shortcut need Cout %% 32 == 0", who);
  B2_ARG_CHECK(res == nullptr || ((uintptr_t)res & 15) == 0, "%s: shortcut must be 16 B aligned", who);
  B2_ARG_CHECK(res == nullptr || Cout % pick_bn(Cout) == 0, "%s: a shortcut needs Cout to fill whole column blocks", who);
  B2_ARG_CHECK(col_sum == nullptr || (o_scale == nullptr && res == nullptr),
               "%s: statistics are taken on the raw output; they do not combine with an output BatchNorm / shortcut", who);
  B2_ARG_CHECK(fin_scale == nullptr || (fin_shift && fin_gamma && fin_beta && fin_counter && col_sum),
               "%s: BatchNorm finalisation needs gamma/beta/shift/counter and the statistics buffers", who);
  ATransform at = {a_scale, a_shift, a_relu};
  EpiParams ep = plain_epi(y, Cout, nullptr, nullptr, 1, relu, col_sum, col_sumsq);
  ep.store = y != nullptr;
  ep.o_scale = o_scale;
  ep.o_shift = o_shift;
  ep.res = (const bf16*)res;
  ep.r_scale = r_scale;
  ep.r_shift = r_shift;
  if (fin_scale != nullptr) {
    const int P = (H + 2 * pad - R) / stride + 1, Q = (W + 2 * pad - S) / stride + 1;
    const double count = (double)Nimg * P * Q;
    ep.fin.scale = fin_scale;
    ep.fin.shift = fin_shift;
    ep.fin.gamma = fin_gamma;
    ep.fin.beta = fin_beta;
    ep.fin.running_mean = fin_running_mean;
    ep.fin.running_var = fin_running_var;
    ep.fin.counter = fin_counter;
    ep.fin.inv_count = (float)(1.0 / count);
    ep.fin.unbias = count > 1 ? (float)(count / (count - 1.0)) : 1.f;
    ep.fin.eps = eps;
    ep.fin.momentum = momentum;
  }
  return conv_common(x, Nimg, H, W, C, w, Cout, R, S, stride, pad, at, ep, (cudaStream_t)stream, who);
}

B2_API int b2_bn_finalize_nhwc(const float* sum, const float* sumsq, const float* gamma, const float* beta,
                               float* running_mean, float* running_var, long count, float eps, float momentum,
                               int train, float* scale, float* shift, int C, void* stream) {
  B2_ARG_CHECK(gamma && beta && scale && shift && C > 0 && count > 0, "b2_bn_finalize_nhwc: null pointer or empty");
  B2_ARG_CHECK(train ? (sum && sumsq) : (running_mean && running_var), "b2_bn_finalize_nhwc: missing statistics");
  const float inv = (float)(1.0 / (double)count);
  const float unbias = count > 1 ? (float)((double)count / (double)(count - 1)) : 1.f;
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(sum, sumsq, gamma, beta, running_mean,
                                                                         running_var, inv, unbias, eps, momentum,
                                                                         train, scale, shift, C);
  B2_LAUNCH_CHECK("bn_finalize_kernel");
  return 0;
}

// BatchNorm statistics + finalisation of a 1x1 convolution's output WITHOUT computing the output (Gram-matrix
// form, see gram_stats_kernel) ; include/b200lrcn.h
// [kGramCopies x (K*K + K)] partial Gram / sums | [4096] per-channel accumulators | [256] channel-block counters |
// [K*K + K] centred covariance + mean
constexpr int kGramMaxCout = 4096;
B2_API long b2_gram_workspace_floats(int C) { return (long)(kGramCopies + 1) * ((long)C * C + C) + kGramMaxCout + 256; }

template <int KC>
int launch_gram(const CUtensorMap& ta, long M, const ATransform& at, float* workspace, cudaStream_t st) {
  using G = GramLayout<KC>;
  static B2PerDeviceOnce attr_set;
  if (attr_set.needed()) {
    B2_CUDA_CHECK(cudaFuncSetAttribute(gram_stats_kernel<KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, G::kSmem));
    attr_set.mark();
  }
  const int rows_per_super = KC == 64 ? 2 * BM : BM;
  const int num_super = b2_ceil_div(M, rows_per_super);
  const int grid = num_super < b2_num_sms() ? num_super : b2_num_sms();
  gram_stats_kernel<KC><<<grid, kGramThreads, G::kSmem, st>>>(ta, (int)M, at, workspace);
  B2_LAUNCH_CHECK("gram_stats_kernel");
  return 0;
}

B2_API int b2_conv1x1_gram_bnstats_bf16(const void* x, long M, int C, const void* w, int Cout, const float* a_scale,
                                        const float* a_shift, int a_relu, float* workspace, float* col_sum,
                                        float* col_sumsq, const float* fin_gamma, const float* fin_beta,
                                        float* fin_running_mean, float* fin_running_var, float* fin_scale,
                                        float* fin_shift, float eps, float momentum, void* stream) {
  const char* who = "b2_conv1x1_gram_bnstats_bf16";
  B2_ARG_CHECK(x && w && a_scale && a_shift && workspace && fin_gamma && fin_beta && fin_scale && fin_shift,
               "%s: null pointer", who);
  B2_ARG_CHECK(M > 0 && M < (1L << 31) && Cout > 0 && Cout <= kGramMaxCout, "%s: empty or oversized shape", who);
  B2_ARG_CHECK(C == 64 || C == 128 || C == 256, "%s: 64, 128 or 256 input channels (got %d)", who, C);
  B2_ARG_CHECK((col_sum == nullptr) == (col_sumsq == nullptr), "%s: col_sum and col_sumsq go together", who);
  B2_ARG_CHECK(((uintptr_t)x & 15) == 0 && ((uintptr_t)workspace & 15) == 0, "%s: x / workspace must be 16 B aligned", who);
  if (int r = load_driver_entry_points()) return r;
  cudaStream_t st = (cudaStream_t)stream;
  CUtensorMap ta;
  if (int r = make_tmap_2d(&ta, x, M, C, C, BM)) return r;
  // one memset: the partial accumulators and the finalisation accumulators / counters are adjacent
  float* fin_acc = workspace + (long)kGramCopies * ((long)C * C + C);
  B2_CUDA_CHECK(cudaMemsetAsync(workspace, 0, ((size_t)kGramCopies * ((size_t)C * C + C) + kGramMaxCout + 256) * sizeof(float), st));
  ATransform at = {a_scale, a_shift, a_relu};
  int r = C == 64 ? launch_gram<64>(ta, M, at, workspace, st)
                  : (C == 128 ? launch_gram<128>(ta, M, at, workspace, st) : launch_gram<256>(ta, M, at, workspace, st));
  if (r) return r;
  BnFinal f = {};
  f.scale = fin_scale;
  f.shift = fin_shift;
  f.gamma = fin_gamma;
  f.beta = fin_beta;
  f.running_mean = fin_running_mean;
  f.running_var = fin_running_var;
  f.inv_count = (float)(1.0 / (double)M);
  f.unbias = M > 1 ? (float)((double)M / ((double)M - 1.0)) : 1.f;
  f.eps = eps;
  f.momentum = momentum;
  float* cov = fin_acc + kGramMaxCout + 256;                           // centred covariance + mean vector
  gram_reduce_kernel<<<b2_ceil_div(C * C / 4, 256), 256, 0, st>>>(workspace, cov, C, f.inv_count);
  B2_LAUNCH_CHECK("gram_reduce_kernel");
  static B2PerDeviceOnce fin_attr;
  const size_t fin_smem = (size_t)(kGramFinRows * 256 + 256 + kGramFinWarps * kGramFinPerWarp * 256) * sizeof(float);
  if (fin_attr.needed()) {
    B2_CUDA_CHECK(cudaFuncSetAttribute(gram_finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fin_smem));
    fin_attr.mark();
  }
  dim3 fgrid(b2_ceil_div(Cout, kGramFinWarps * kGramFinPerWarp), b2_ceil_div(C, kGramFinRows));
  gram_finalize_kernel<<<fgrid, 32 * kGramFinWarps, (size_t)(kGramFinRows * C + C + kGramFinWarps * kGramFinPerWarp * C) * sizeof(float), st>>>(
      cov, (const bf16*)w, Cout, C, fin_acc, reinterpret_cast<unsigned int*>(fin_acc + kGramMaxCout), col_sum, col_sumsq, f,
      (float)M);
  B2_LAUNCH_CHECK("gram_finalize_kernel");
  return 0;
}
