// tcgen05 / TMEM / TMA GEMM and implicit-GEMM convolution for sm_100a.
//
//   D[M,N] = A[M,K] * B[N,K]^T (+ bias[N]) (ReLU) ; fp32 accumulate in TMEM, bf16 operands.
//
// A is fed by TMA either from a plain row-major matrix (2-D tiled tensor map: Linear layers,
// 1x1 stride-1 convolutions over NHWC activations, hoisted LSTM gate GEMMs) or straight from an
// NHWC activation tensor through an IM2COL-mode tensor map (3x3 / strided convolutions: one TMA
// per filter tap and 64-channel slab, zero padding and stride handled by the copy engine), so no
// im2col matrix ever exists in HBM.  B (the weights, [N,K] K-major) always comes from a 2-D map.
//
// Kernel anatomy: persistent CTAs (one per SM), 6 warps:
//   warp 0      TMA producer           (smem ring of kStages {A 128x64, B BNx64} bf16 tiles, SW128)
//   warp 1      MMA issuer + TMEM owner (tcgen05.mma cta_group::1, M=128, N=BN, K=16 per instr)
//   warps 2..5  epilogue               (tcgen05.ld 32x32b -> bias/ReLU/bf16 -> global, plus the
//                                       per-column sum / sum-of-squares the train-mode BatchNorm
//                                       that follows needs, so BN statistics cost no extra pass)
// TMEM holds two BN-column accumulators so the epilogue of tile i overlaps the MMAs of tile i+1.
#include "tc_ptx.cuh"

#include <cudaTypedefs.h>
#include <mutex>

namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 B = one SWIZZLE_128B row
constexpr int UMMA_K = 16;
constexpr int kEpiWarps = 8;                      // two warps per TMEM lane quarter
constexpr int kThreads = 64 + 32 * kEpiWarps;    // TMA warp + MMA warp + epilogue warps

using namespace tc;

struct ConvGeom {   // im2col producer geometry (all zero for plain GEMM)
  int is_conv;
  int P, Q;         // output height / width
  int S;            // filter width (taps per filter row)
  int c_slabs;      // C / 64
  int stride;
  int lower_w, lower_h;   // = -pad
};

struct EpiParams {
  void* D;
  long ldd;          // elements
  const float* bias; // [N] or null
  const float* bias2; // [N] or null (second bias vector, e.g. LSTM b_hh)
  float* col_sum;    // [N] or null  (atomicAdd)
  float* col_sumsq;  // [N] or null
  int out_bf16;      // 1: bf16 out, 0: fp32 out
  int relu;
  int tma_store;     // bf16 output leaves through smem staging + TMA bulk stores (tmap_d valid)
};

constexpr int kStgBytes = 32 * 64;  // epilogue staging tile: 32 rows x 32 bf16 (64-byte rows, SWIZZLE_64B)

template <int BN>
struct SmemLayout {
  static constexpr int kStageBytes = (BM * BK + BN * BK) * 2;
  static constexpr int kStages = (BN >= 256) ? 4 : (BN >= 128 ? 5 : (BN >= 64 ? 7 : 8));
  static constexpr int kStgBufs = (BN >= 256) ? 1 : 2;   // staging tiles per epilogue warp
  static constexpr int kTileBytes = kStageBytes * kStages;
  static constexpr int kBarOffset = kTileBytes;
  static constexpr int kStatOffset = kBarOffset + (2 * kStages + 4) * 8 + 16;     // per-CTA column statistics [4 quarters][2][BN]
  static constexpr int kScratchOffset = kStatOffset + 4 * 2 * BN * 4;
  static constexpr int kScratchOffset1k = (kScratchOffset + 1023) / 1024 * 1024;        // staging tiles 1024-aligned
  static constexpr int kTotal = kScratchOffset1k + kEpiWarps * kStgBufs * kStgBytes + 1024;   // +1024 alignment slack
};

template <int BN>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const __grid_constant__ CUtensorMap tmap_d, int M, int N, int K, ConvGeom g, EpiParams ep) {
  using L = SmemLayout<BN>;
  constexpr int kStages = L::kStages;
  constexpr uint32_t kTmemCols = 2 * BN;  // two accumulators (32 <= cols <= 512, power of two)
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte alignment (SWIZZLE_128B) computed as an OFFSET so that every derived pointer keeps its
  // shared-memory provenance (integer round-tripping made the compiler emit generic LD/ST/atomics).
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tfull_bar = empty_bar + kStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  // Column statistics (sum, sum of squares of the stored values) are accumulated per CTA in shared
  // memory over all consecutive tiles of one n-block and flushed to global memory only when the
  // n-block changes (tiles are ordered m-fastest, so that is at most n_blocks times per CTA).
  // One private copy per TMEM lane quarter: the two warps of a quarter own disjoint chunks, so a slot
  // has exactly one writer and plain read-modify-write replaces shared-memory CAS atomics.
  float* stat_s = reinterpret_cast<float*>(smem + L::kStatOffset);   // [4][2][BN]
  uint8_t* staging_s = smem + L::kScratchOffset1k;
  const bool want_stats = ep.col_sum != nullptr;
  for (int i = threadIdx.x; i < 8 * BN; i += kThreads) stat_s[i] = 0.f;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_blocks = (M + BM - 1) / BM;
  const int n_blocks = (N + BN - 1) / BN;
  const int num_tiles = m_blocks * n_blocks;
  const int k_blocks = (K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], kEpiWarps);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tc_alloc(tmem_slot, kTmemCols);
    tc_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int n_blk = tile / m_blocks;           // m-fastest: concurrent CTAs share one weight tile
        const int m_blk = tile - n_blk * m_blocks;
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
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * L::kStageBytes;
          uint8_t* sb = sa + BM * BK * 2;
          mbar_expect_tx(&full_bar[stage], L::kStageBytes);
          if (g.is_conv) {
            const int tap = kb / g.c_slabs;
            const int c0 = (kb - tap * g.c_slabs) * BK;
            const int r = tap / g.S;
            const int s = tap - r * g.S;
            tma_load_im2col_4d(sa, &tmap_a, &full_bar[stage], c0, cw, ch, cn, (uint16_t)s, (uint16_t)r);
          } else {
            tma_load_2d(sa, &tmap_a, &full_bar[stage], kb * BK, m_blk * BM);
          }
          tma_load_2d(sb, &tmap_b, &full_bar[stage], kb * BK, n_blk * BN);
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    constexpr uint32_t idesc = make_idesc(BM, BN);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
      for (int kb = 0; kb < k_blocks; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = smem_u32(smem + stage * L::kStageBytes);
          const uint32_t sb = sa + BM * BK * 2;
          const uint64_t da = make_sw128_desc(sa);
          const uint64_t db = make_sw128_desc(sb);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // +32 B per K=16 step inside the 128 B swizzle row (encoded >>4 -> +2)
            tc_mma_bf16(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) != 0);
          }
          tc_commit(&empty_bar[stage]);                       // frees the smem slot when the MMAs retire
          if (kb == k_blocks - 1) tc_commit(&tfull_bar[acc]);  // accumulator complete -> epilogue
        }
        __syncwarp();
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else {
    // =========================== epilogue (warps 2..9) ===========================
    const int quarter = warp & 3;          // TMEM lane quarter this warp may read (hardware rule: warp % 4)
    const int half = (warp - 2) >> 2;      // the two warps of a quarter take alternate 32-column chunks
    const int epi_tid = threadIdx.x - 64;
    uint8_t* stg_base = staging_s + (warp - 2) * L::kStgBufs * kStgBytes;
    int stg_buf = 0;
    int it = 0;
    int stat_nblk = -1;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int n_blk = tile / m_blocks;
      const int m_blk = tile - n_blk * m_blocks;
      if (want_stats && n_blk != stat_nblk) {
        if (stat_nblk >= 0) {                       // flush the finished n-block (rare)
          asm volatile("bar.sync 1, 256;" ::: "memory");
          for (int c = epi_tid; c < BN; c += 32 * kEpiWarps) {
            const int col = stat_nblk * BN + c;
            if (col < N) {
              atomicAdd(ep.col_sum + col, stat_s[c] + stat_s[2 * BN + c] + stat_s[4 * BN + c] + stat_s[6 * BN + c]);
              atomicAdd(ep.col_sumsq + col,
                        stat_s[BN + c] + stat_s[3 * BN + c] + stat_s[5 * BN + c] + stat_s[7 * BN + c]);
            }
#pragma unroll
            for (int qq = 0; qq < 8; ++qq) stat_s[qq * BN + c] = 0.f;
          }
          asm volatile("bar.sync 1, 256;" ::: "memory");
        }
        stat_nblk = n_blk;
      }
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const int row = m_blk * BM + quarter * 32 + lane;
      const bool row_ok = row < M;
#pragma unroll 1
      for (int ch = half; ch < BN / 32; ch += 2) {
        const int col0 = n_blk * BN + ch * 32;
        if (col0 >= N) break;  // warp-uniform
        uint32_t raw[32];
        tc_ld32(tmem_base + (uint32_t)(acc * BN + ch * 32) + ((uint32_t)(quarter * 32) << 16), raw);
        tc_wait_ld();
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]);
        const bool full_chunk = (col0 + 32 <= N);
        // every option below is a WARP-UNIFORM branch around its own loop: predicating 64 bias loads
        // and adds per chunk (the first version) cost more issue slots than the whole rest of the epilogue
        if (ep.bias != nullptr) {
          if (full_chunk) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 bv = __ldg(reinterpret_cast<const float4*>(ep.bias + col0 + j));
              v[j] += bv.x; v[j + 1] += bv.y; v[j + 2] += bv.z; v[j + 3] += bv.w;
            }
          } else {
            for (int j = 0; j < 32; ++j)
              if (col0 + j < N) v[j] += __ldg(ep.bias + col0 + j);
          }
        }
        if (ep.bias2 != nullptr) {
          for (int j = 0; j < 32; ++j)
            if (col0 + j < N) v[j] += __ldg(ep.bias2 + col0 + j);
        }
        if (ep.relu) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.0f);
        }
        if (ep.out_bf16) {
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) pk[j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
          uint8_t* stg = stg_base + stg_buf * kStgBytes;
          const bool staged = ep.tma_store || want_stats;
          if (staged) {
            // The 32x32 bf16 chunk goes to a 64-byte-swizzled staging tile (lane = row, 4 x STS.128,
            // conflict-free).  From there (a) one TMA bulk store writes it to global memory in full
            // 64-byte row segments without occupying the LSU or any registers, and (b) the BN
            // statistics are column sums read straight back from the tile (lane pairs share a word).
            if (ep.tma_store && lane == 0) bulk_wait_read<L::kStgBufs - 1>();   // tile free again?
            __syncwarp();
            if (want_stats && !row_ok) {
#pragma unroll
              for (int j = 0; j < 16; ++j) pk[j] = 0u;       // rows past M: no statistics (TMA clips the store)
            }
            const int sw = (lane >> 1) & 3;
#pragma unroll
            for (int c = 0; c < 4; ++c)
              *reinterpret_cast<uint4*>(stg + lane * 64 + ((c ^ sw) << 4)) =
                  make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
          }
          if (ep.tma_store) {
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&tmap_d, stg, col0, m_blk * BM + quarter * 32);
              bulk_commit();
            }
          } else if (row_ok) {
            if (staged) __syncwarp();
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
              for (int j = 0; j < 16; j += 4)
                *reinterpret_cast<uint4*>(dp + 2 * j) = make_uint4(pk[j], pk[j + 1], pk[j + 2], pk[j + 3]);
            } else {
              for (int j = 0; j < 32; ++j)
                if (col0 + j < N) dp[j] = __float2bfloat16_rn(v[j]);
            }
          } else if (staged) {
            __syncwarp();
          }
          if (want_stats) {
            float s1a = 0.f, s1b = 0.f, s1c = 0.f, s1d = 0.f, s2a = 0.f, s2b = 0.f, s2c = 0.f, s2d = 0.f;
            const int w = lane >> 1;                       // bf16x2 word of the row this lane pair reads
            const int sh = (lane & 1) ? 0 : 16;            // even lanes take the low bf16 of the word
            const uint8_t* colp = stg + ((w & 3) << 2);
            const int wc = w >> 2;
#pragma unroll
            for (int r = 0; r < 32; r += 4) {              // four independent accumulation chains
              uint32_t u[4];
#pragma unroll
              for (int q4 = 0; q4 < 4; ++q4)
                u[q4] = *reinterpret_cast<const uint32_t*>(colp + (r + q4) * 64 + ((wc ^ (((r + q4) >> 1) & 3)) << 4));
              const float x0 = __uint_as_float((u[0] << sh) & 0xffff0000u);
              const float x1 = __uint_as_float((u[1] << sh) & 0xffff0000u);
              const float x2 = __uint_as_float((u[2] << sh) & 0xffff0000u);
              const float x3 = __uint_as_float((u[3] << sh) & 0xffff0000u);
              s1a += x0; s1b += x1; s1c += x2; s1d += x3;
              s2a = fmaf(x0, x0, s2a); s2b = fmaf(x1, x1, s2b); s2c = fmaf(x2, x2, s2c); s2d = fmaf(x3, x3, s2d);
            }
            float* st = stat_s + quarter * 2 * BN + ch * 32 + lane;
            st[0] += (s1a + s1b) + (s1c + s1d);
            st[BN] += (s2a + s2b) + (s2c + s2d);
          }
          if (staged) stg_buf = (stg_buf + 1) % L::kStgBufs;
        } else if (row_ok) {
          float* dp = reinterpret_cast<float*>(ep.D) + (long)row * ep.ldd + col0;
          if (full_chunk && ((reinterpret_cast<uintptr_t>(dp) & 31) == 0)) {
#pragma unroll
            for (int j = 0; j < 32; j += 8)
              asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dp + j), "f"(v[j]), "f"(v[j + 1]),
                           "f"(v[j + 2]), "f"(v[j + 3]), "f"(v[j + 4]), "f"(v[j + 5]), "f"(v[j + 6]), "f"(v[j + 7])
                           : "memory");
          } else if (full_chunk && ((reinterpret_cast<uintptr_t>(dp) & 15) == 0)) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(dp + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          } else {
            for (int j = 0; j < 32; ++j)
              if (col0 + j < N) dp[j] = v[j];
          }
        }
        if (want_stats && !ep.out_bf16) {
          // fp32 output: column sums over this warp's 32 rows by butterfly reduce-scatter (lane l ends with column l)
          float s2[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            v[j] = row_ok ? v[j] : 0.0f;
            s2[j] = v[j] * v[j];
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
          float* st = stat_s + quarter * 2 * BN + ch * 32 + lane;
          st[0] += v[0];
          st[BN] += s2[0];
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
    }
    if (ep.tma_store && lane == 0) bulk_wait_all();   // all bulk stores of this warp have landed
  }

  tc_fence_before();
  __syncthreads();
  if (want_stats) {   // flush the last n-block this CTA worked on
    int last_tile = -1;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) last_tile = tile;
    if (last_tile >= 0) {
      const int nb = last_tile / m_blocks;
      for (int c = threadIdx.x; c < BN; c += kThreads) {
        const int col = nb * BN + c;
        if (col < N) {
          atomicAdd(ep.col_sum + col, stat_s[c] + stat_s[2 * BN + c] + stat_s[4 * BN + c] + stat_s[6 * BN + c]);
          atomicAdd(ep.col_sumsq + col, stat_s[BN + c] + stat_s[3 * BN + c] + stat_s[5 * BN + c] + stat_s[7 * BN + c]);
        }
      }
    }
  }
  if (warp == 1) {
    tc_fence_after();
    tc_dealloc(tmem_base, kTmemCols);
  }
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

template <int BN>
int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& td, int M, int N, int K,
                const ConvGeom& g, const EpiParams& ep, cudaStream_t stream) {
  using L = SmemLayout<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    B2_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal));
    attr_set = true;
  }
  const int tiles = b2_ceil_div(M, BM) * b2_ceil_div(N, BN);
  const int grid = tiles < b2_num_sms() ? tiles : b2_num_sms();
  gemm_tc_kernel<BN><<<grid, kThreads, L::kTotal, stream>>>(ta, tb, td, M, N, K, g, ep);
  B2_LAUNCH_CHECK("gemm_tc_kernel");
  return 0;
}

int pick_bn(int N) {
  if (N > 128) return 256;
  if (N > 64) return 128;
  if (N > 32) return 64;
  return 32;
}

int dispatch(int bn, const CUtensorMap& ta, const CUtensorMap& tb, int M, int N, int K, const ConvGeom& g,
             EpiParams ep, cudaStream_t stream) {
  // bf16 outputs whose rows are 16-byte multiples leave through TMA bulk stores
  CUtensorMap td = ta;
  ep.tma_store = 0;
  if (ep.out_bf16 && (ep.ldd % 8) == 0 && ((uintptr_t)ep.D & 15) == 0) {
    if (int r = make_tmap_2d(&td, ep.D, M, N, ep.ldd, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B)) return r;
    ep.tma_store = 1;
  }
  switch (bn) {
    case 256: return launch_gemm<256>(ta, tb, td, M, N, K, g, ep, stream);
    case 128: return launch_gemm<128>(ta, tb, td, M, N, K, g, ep, stream);
    case 64: return launch_gemm<64>(ta, tb, td, M, N, K, g, ep, stream);
    default: return launch_gemm<32>(ta, tb, td, M, N, K, g, ep, stream);
  }
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
  const int bn = pick_bn(N);
  CUtensorMap ta, tb;
  if (int r = make_tmap_2d(&ta, A, M, K, lda, BM)) return r;
  if (int r = make_tmap_2d(&tb, B, N, K, ldb, bn)) return r;
  ConvGeom g = {};
  EpiParams ep = {D, ldd, bias, bias2, col_sum, col_sumsq, out_bf16, relu, 0};
  return dispatch(bn, ta, tb, M, N, K, g, ep, (cudaStream_t)stream);
}

// y[N,P,Q,Cout] = conv(x[N,H,W,C], w[Cout,R,S,C]) ; NHWC bf16, C % 64 == 0 ; see include/b200lrcn.h
B2_API int b2_conv2d_nhwc_bf16(const void* x, int Nimg, int H, int W, int C, const void* w, int Cout, int R, int S,
                               int stride, int pad, void* y, const float* bias, int out_bf16, int relu,
                               float* col_sum, float* col_sumsq, void* stream) {
  B2_ARG_CHECK(x && w && y && Nimg > 0 && H > 0 && W > 0 && Cout > 0, "b2_conv2d_nhwc_bf16: null pointer or empty shape");
  B2_ARG_CHECK(C % 64 == 0, "b2_conv2d_nhwc_bf16: C must be a multiple of 64 (got %d)", C);
  B2_ARG_CHECK(R >= 1 && S >= 1 && R <= 7 && S <= 7 && stride >= 1 && stride <= 8 && pad >= 0 && pad <= 3,
               "b2_conv2d_nhwc_bf16: unsupported filter geometry R=%d S=%d stride=%d pad=%d", R, S, stride, pad);
  B2_ARG_CHECK((col_sum == nullptr) == (col_sumsq == nullptr), "b2_conv2d_nhwc_bf16: col_sum and col_sumsq go together");
  const int P = (H + 2 * pad - R) / stride + 1;
  const int Q = (W + 2 * pad - S) / stride + 1;
  B2_ARG_CHECK(P > 0 && Q > 0, "b2_conv2d_nhwc_bf16: empty output");
  if (int r = load_driver_entry_points()) return r;
  const long Ml = (long)Nimg * P * Q;
  B2_ARG_CHECK(Ml < (1L << 31), "b2_conv2d_nhwc_bf16: too many output pixels");
  const int M = (int)Ml;
  const int K = R * S * C;
  const int bn = pick_bn(Cout);
  EpiParams ep = {y, (long)Cout, bias, nullptr, col_sum, col_sumsq, out_bf16, relu, 0};
  CUtensorMap ta, tb;
  if (int r = make_tmap_2d(&tb, w, Cout, K, K, bn)) return r;
  if (R == 1 && S == 1 && stride == 1 && pad == 0) {
    if (int r = make_tmap_2d(&ta, x, M, C, C, BM)) return r;
    ConvGeom g = {};
    return dispatch(bn, ta, tb, M, Cout, K, g, ep, (cudaStream_t)stream);
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
  ConvGeom g = {1, P, Q, S, C / 64, stride, -pad, -pad};
  return dispatch(bn, ta, tb, M, Cout, K, g, ep, (cudaStream_t)stream);
}
