// bf16 tensor-core path of the small TimeDistributed CNN (notebook LRCN nb:148-193, lrcn/backup_ucf50.py:105-151;
// BASELINE.json configs[0]): conv 3->16->32->64, 3x3 / pad 1, train-mode BatchNorm, ReLU, 2x2 max-pool.
//
// Channel counts of 16 / 32 / 64 are too narrow for the 64-channel SWIZZLE_128B tiles of gemm_tc.cu / conv_halo.cu, so
// this file has its own halo-tile kernels in which ONE PIXEL IS ONE SWIZZLE ROW of 2*C bytes:
//     C = 16 -> 32-byte rows, SWIZZLE_32B      C = 32 -> 64-byte rows, SWIZZLE_64B      C = 64 -> 128-byte rows, SWIZZLE_128B
// Activations are NHWC bf16.  A tile = TH whole output rows of one image; ONE TMA (4-D box {C, W+2, TH+2, 1}, out-of-image
// coordinates zero-filled = the conv padding) brings its halo into shared memory, and every filter tap (r, s) reads that
// SAME tile through a descriptor whose start address is shifted by (r*(W+2) + s) pixels ("padded-width" row indexing: the
// two extra columns per row are computed and discarded).  The tensor core applies the swizzle to absolute shared-memory
// address bits, exactly as TMA did when it wrote the tile, so shifted starts need no re-layout.
//
//   sc_conv_kernel<CIN, COUT>   forward conv AND data gradient (the same kernel on the flipped / transposed filter):
//                               A = halo tile, K-major (pixel rows x CIN), one K = 16 MMA per tap and 16-channel slab;
//                               NB = ceil(TH*(W+2) / 128) accumulators of 128 x COUT per tile in TMEM, double buffered;
//                               the 9 x [COUT x CIN] weight tiles stay resident; epilogue: + bias, raw bf16 NHWC rows,
//                               per-channel sum / sum of squares of the stored values for the following BatchNorm.
//   sc_wgrad_kernel<CIN, COUT>  dW[co][r][s][ci] = sum_pixels dz[p][co] * x[p + (r, s)][ci]: BOTH operands MN-major (the
//                               reduction runs over pixels = swizzle rows).  M = 128 is filled with 128 / CIN copies of the
//                               x tile shifted by one pixel each (leading-dimension byte offset = one row): rows
//                               [s*CIN, (s+1)*CIN) of the accumulator are filter column s, so ONE MMA per filter row r and
//                               16-pixel step produces three taps; the three fp32 accumulators live in TMEM for the whole
//                               kernel and are drained once with atomics.
//   sc_conv1_*                  the 3-channel first layer (K = 27) stays on CUDA cores: fp32 NCHW frames in, bf16 NHWC out.
//   sc_act_pool_*               BatchNorm + ReLU (+ 2x2 max-pool) forward / backward (first-max routing, two-pass BN
//                               backward) on NHWC bf16, 128-bit accesses;  sc_nhwc_to_chw_* the channel-major flatten the
//                               LSTM's W_ih columns expect (nb:186) fused with the feature dropout.
#include "tc_ptx.cuh"

#include <cudaTypedefs.h>
#include <mutex>

namespace {
using namespace tc;

constexpr int kTaps = 9;

struct ScGeom {
  int N, H, W;
  int TH, Wp, NB, Mt;          // output rows per tile, padded width, 128-row accumulator blocks, TH * Wp
  int tiles_per_img, num_tiles;
  int halo_px;                 // (TH + 2) * Wp
  int stage_bytes, stages;
  uint32_t magic_wp, magic_tpi;
};

__device__ __forceinline__ void tma_load_4d(void* dst, const void* tmap, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
// layout-type field of a shared-memory matrix descriptor for rows of RB bytes
template <int RB>
__device__ __forceinline__ constexpr uint64_t swz_bits() {
  return (uint64_t)(RB == 128 ? 2 : (RB == 64 ? 4 : 6)) << 61;
}
// K-major: rows of RB bytes (= the whole K extent of one tap), 8-row groups 8*RB bytes apart
template <int RB>
__device__ __forceinline__ uint64_t kmajor_desc(uint32_t addr) {
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)((8 * RB) >> 4) << 32) | ((uint64_t)1 << 46) |
         swz_bits<RB>();
}
// MN-major: RB/2 contiguous MN elements per row (one row per K index), 8-row K groups 8*RB bytes apart, MN blocks `lbo`
// bytes apart
template <int RB>
__device__ __forceinline__ uint64_t mnmajor_desc(uint32_t addr, uint32_t lbo) {
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((8 * RB) >> 4) << 32) |
         ((uint64_t)1 << 46) | swz_bits<RB>();
}

// column sums of a 32 x 32 tile held as v[j] = tile[lane][j]: after the butterfly lane l holds sum_rows tile[.][l]
__device__ __forceinline__ float colsum32(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = up ? v[i] : v[i + off];
      const float keep = up ? v[i + off] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

// ------------------------------------------------------------------------------------------------------------------
// forward conv / data gradient on the tensor core
// ------------------------------------------------------------------------------------------------------------------
constexpr int kConvThreads = 64 + 32 * 8;     // TMA warp, MMA warp, 8 epilogue warps

template <int CIN, int COUT>
__global__ void __launch_bounds__(kConvThreads, 1)
sc_conv_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
               bf16* __restrict__ y, ScGeom g, const float* __restrict__ bias, float* col_sum, float* col_sumsq) {
  constexpr int RB = CIN * 2;                          // bytes per pixel row
  constexpr int KS = CIN / 16;                         // K = 16 steps per tap
  constexpr int kWBytes = COUT * RB;                   // one tap's [COUT x CIN] weight tile ...
  constexpr int kWTile = (kWBytes + 1023) / 1024 * 1024;   // ... in a slot that keeps the 1024-byte stage alignment
  constexpr int CW = COUT < 32 ? COUT : 32;            // columns per TMEM load
  constexpr int kChunks = COUT / CW;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* w_s = smem;
  uint8_t* stage_s = smem + kTaps * kWTile;
  uint8_t* after = stage_s + g.stages * g.stage_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(after);
  uint64_t* empty_bar = full_bar + 4;
  uint64_t* tfull_bar = empty_bar + 4;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* w_bar = tempty_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_bar + 1);
  float* stat_s = reinterpret_cast<float*>(after + 128);          // [8 warps][2][COUT]
  uint32_t* xpose_s = reinterpret_cast<uint32_t*>(after + 128 + 8 * 2 * COUT * 4);   // [8 warps][16][33] words
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const bool want_stats = col_sum != nullptr;
  uint32_t tmem_cols = 32;
  while (tmem_cols < (uint32_t)(2 * g.NB * COUT)) tmem_cols <<= 1;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_x);
    prefetch_tmap(&tmap_w);
    for (int i = 0; i < g.stages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 8);
    }
    mbar_init(w_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tc_alloc(tmem_slot, tmem_cols);
    tc_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(w_bar, kTaps * kWBytes);
      for (int t = 0; t < kTaps; ++t) tma_load_2d(w_s + t * kWTile, &tmap_w, w_bar, t * CIN, 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < g.num_tiles; tile += gridDim.x) {
        const int n = g.tiles_per_img == 1 ? tile : (int)__umulhi((uint32_t)tile, g.magic_tpi);
        const int h0 = (tile - n * g.tiles_per_img) * g.TH;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        mbar_expect_tx(&full_bar[stage], (uint32_t)(g.halo_px * RB));
        tma_load_4d(stage_s + stage * g.stage_bytes, &tmap_x, &full_bar[stage], 0, -1, h0 - 1, n);
        if (++stage == g.stages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc(128, COUT);
    mbar_wait(w_bar, 0);
    const uint64_t db0 = kmajor_desc<RB>(smem_u32(w_s));
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < g.num_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      mbar_wait(&tempty_bar[acc], ((it >> 1) & 1) ^ 1);
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t d0 = tmem_base + (uint32_t)(acc * g.NB * COUT);
        const uint64_t da0 = kmajor_desc<RB>(smem_u32(stage_s + stage * g.stage_bytes));
        for (int b = 0; b < g.NB; ++b) {
#pragma unroll
          for (int t = 0; t < kTaps; ++t) {
            const uint32_t a_off = (uint32_t)((b * 128 + (t / 3) * g.Wp + (t % 3)) * (RB / 16));
#pragma unroll
            for (int k = 0; k < KS; ++k)
              tc_mma_bf16(d0 + (uint32_t)(b * COUT), da0 + (uint64_t)(a_off + k * 2),
                          db0 + (uint64_t)(t * (kWTile / 16) + k * 2), idesc, (t | k) != 0);
          }
        }
        tc_commit(&empty_bar[stage]);
        tc_commit(&tfull_bar[acc]);
      }
      __syncwarp();
      if (++stage == g.stages) {
        stage = 0;
        phase ^= 1;
      }
    }
  } else {
    // =========================== epilogue (warps 2..9) ===========================
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;          // the two warps of a lane quarter take alternate accumulator blocks
    float2 acc1[kChunks], acc2[kChunks];       // lane (w, grp): running sum / sum of squares of channels chunk*CW + 2w, 2w+1
#pragma unroll
    for (int c = 0; c < kChunks; ++c) acc1[c] = acc2[c] = make_float2(0.f, 0.f);
    int it = 0;
    for (int tile = blockIdx.x; tile < g.num_tiles; tile += gridDim.x, ++it) {
      const int n = g.tiles_per_img == 1 ? tile : (int)__umulhi((uint32_t)tile, g.magic_tpi);
      const int h0 = (tile - n * g.tiles_per_img) * g.TH;
      const int acc = it & 1;
      mbar_wait(&tfull_bar[acc], (it >> 1) & 1);
      tc_fence_after();
      for (int b = half; b < g.NB; b += 2) {
        const int u = b * 128 + quarter * 32 + lane;
        const int pl = (int)__umulhi((uint32_t)u, g.magic_wp);
        const int q = u - pl * g.Wp;
        const bool row_ok = u < g.Mt && q < g.W && h0 + pl < g.H;
        bf16* dp = y + (((long)n * g.H + h0 + pl) * g.W + q) * COUT;
#pragma unroll
        for (int c = 0; c < kChunks; ++c) {
          uint32_t raw[32];
          const uint32_t ta = tmem_base + (uint32_t)(acc * g.NB * COUT + b * COUT + c * CW) + ((uint32_t)(quarter * 32) << 16);
          if (CW == 32) tc_ld32(ta, raw);
          else tc_ld16(ta, raw);
          tc_wait_ld();
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = j < CW ? __uint_as_float(raw[j]) : 0.f;
          if (bias != nullptr) {
#pragma unroll
            for (int j = 0; j < CW; ++j) v[j] += __ldg(bias + c * CW + j);
          }
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < CW / 2; ++j) pk[j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
          if (row_ok) {
#pragma unroll
            for (int s8 = 0; s8 < CW / 16; ++s8)
              asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dp + c * CW + s8 * 16), "r"(pk[8 * s8]),
                           "r"(pk[8 * s8 + 1]), "r"(pk[8 * s8 + 2]), "r"(pk[8 * s8 + 3]), "r"(pk[8 * s8 + 4]),
                           "r"(pk[8 * s8 + 5]), "r"(pk[8 * s8 + 6]), "r"(pk[8 * s8 + 7])
                           : "memory");
          }
          if (want_stats) {
            // statistics of the values as stored (bf16-rounded): the warp's 32 x CW tile goes through shared memory
            // column-pair-major (word j of row `lane` at [j][lane], pitch 33 words: conflict-free both ways); lane
            // (w, grp) then sums column pair w over its CW/2-th share of the rows -- no shuffles (a butterfly
            // reduce-scatter made the epilogue the bottleneck: 62 SHFL per chunk against one SHFL / clk / SM)
            constexpr int kWords = CW / 2;                 // packed column pairs per row
            constexpr int kRowsPer = kWords;               // 32 lanes = kWords column pairs x (32 / kWords) row groups
            uint32_t* xp = xpose_s + (warp - 2) * (16 * 33);
            __syncwarp();
#pragma unroll
            for (int j = 0; j < kWords; ++j) xp[j * 33 + lane] = row_ok ? pk[j] : 0u;
            __syncwarp();
            const int w2 = lane % kWords, grp = lane / kWords;
            float2 s1 = make_float2(0.f, 0.f), s2 = s1;
#pragma unroll
            for (int i = 0; i < kRowsPer; ++i) {
              const uint32_t u = xp[w2 * 33 + grp * kRowsPer + i];
              const float2 xv = make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
              s1 = __fadd2_rn(s1, xv);
              s2 = __ffma2_rn(xv, xv, s2);
            }
            acc1[c] = __fadd2_rn(acc1[c], s1);
            acc2[c] = __fadd2_rn(acc2[c], s2);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
    }
    if (want_stats) {
      constexpr int kWords = CW / 2;
#pragma unroll
      for (int c = 0; c < kChunks; ++c) {        // combine the row groups of a column pair, then one writer per slot
        float2 a = acc1[c], b = acc2[c];
#pragma unroll
        for (int off = kWords; off < 32; off <<= 1) {
          a.x += __shfl_xor_sync(0xffffffffu, a.x, off);
          a.y += __shfl_xor_sync(0xffffffffu, a.y, off);
          b.x += __shfl_xor_sync(0xffffffffu, b.x, off);
          b.y += __shfl_xor_sync(0xffffffffu, b.y, off);
        }
        if (lane < kWords) {
          float* d1 = stat_s + ((warp - 2) * 2 + 0) * COUT + c * CW + 2 * lane;
          float* d2 = stat_s + ((warp - 2) * 2 + 1) * COUT + c * CW + 2 * lane;
          d1[0] = a.x; d1[1] = a.y;
          d2[0] = b.x; d2[1] = b.y;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (want_stats) {
    for (int c = threadIdx.x; c < COUT; c += kConvThreads) {
      float a1 = 0.f, a2 = 0.f;
#pragma unroll
      for (int w8 = 0; w8 < 8; ++w8) {
        a1 += stat_s[(w8 * 2 + 0) * COUT + c];
        a2 += stat_s[(w8 * 2 + 1) * COUT + c];
      }
      atomicAdd(col_sum + c, a1);
      atomicAdd(col_sumsq + c, a2);
    }
  }
  if (warp == 1) {
    tc_fence_after();
    tc_dealloc(tmem_base, tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// weight gradient on the tensor core
// ------------------------------------------------------------------------------------------------------------------
constexpr int kWgThreads = 64 + 32 * 4;       // TMA warp, MMA warp, 4 drain warps
constexpr int kWgStages = 3;

struct WgGeomSc {
  int N, H, W, TH, Wp, Mt, ksteps;            // ksteps = ceil(Mt / 16)
  int tiles_per_img, num_tiles;
  int halo_px;
  int x_bytes, z_bytes;                       // per-stage regions (multiples of 1024)
  uint32_t magic_tpi;
};

template <int CIN, int COUT>
__global__ void __launch_bounds__(kWgThreads, 1)
sc_wgrad_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_z, WgGeomSc g,
                float* __restrict__ dw) {
  constexpr int RBX = CIN * 2, RBZ = COUT * 2;
  constexpr int kAtoms = 128 / CIN;            // copies of the x tile, one pixel apart: filter columns s = 0 .. kAtoms-1
  static_assert(kAtoms >= 3, "three filter columns must fit the 128 accumulator rows");
  constexpr uint32_t kCols = 3 * COUT <= 128 ? 128 : 256;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int stage_bytes = g.x_bytes + g.z_bytes;
  uint8_t* after = smem + kWgStages * stage_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(after);
  uint64_t* empty_bar = full_bar + kWgStages;
  uint64_t* done_bar = empty_bar + kWgStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done_bar + 1);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // the bytes TMA never writes (row padding of the last 16-pixel step, the reach of the shifted taps past the halo box)
  // are multiplied by real data: they must be finite -> zero the stages once
  for (int i = threadIdx.x; i < kWgStages * stage_bytes / 16; i += kWgThreads)
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_x);
    prefetch_tmap(&tmap_z);
    for (int i = 0; i < kWgStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(done_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tc_alloc(tmem_slot, kCols);
    tc_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < g.num_tiles; tile += gridDim.x) {
        const int n = g.tiles_per_img == 1 ? tile : (int)__umulhi((uint32_t)tile, g.magic_tpi);
        const int h0 = (tile - n * g.tiles_per_img) * g.TH;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sx = smem + stage * stage_bytes;
        mbar_expect_tx(&full_bar[stage], (uint32_t)(g.halo_px * RBX + g.Mt * RBZ));
        tma_load_4d(sx, &tmap_x, &full_bar[stage], 0, -1, h0 - 1, n);
        tma_load_4d(sx + g.x_bytes, &tmap_z, &full_bar[stage], 0, 0, h0, n);    // columns >= W, rows >= H: zero filled
        if (++stage == kWgStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc(128, COUT) | (1u << 15) | (1u << 16);     // both operands MN-major
    int stage = 0;
    uint32_t phase = 0;
    bool first = true;
    for (int tile = blockIdx.x; tile < g.num_tiles; tile += gridDim.x) {
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t sx = smem_u32(smem + stage * stage_bytes);
        const uint32_t sz = sx + (uint32_t)g.x_bytes;
        for (int j = 0; j < g.ksteps; ++j) {
          const uint64_t dz = mnmajor_desc<RBZ>(sz + (uint32_t)(j * 16 * RBZ), 8 * RBZ);
#pragma unroll
          for (int r = 0; r < 3; ++r) {
            const uint64_t dx = mnmajor_desc<RBX>(sx + (uint32_t)((j * 16 + r * g.Wp) * RBX), RBX);
            tc_mma_bf16(tmem_base + (uint32_t)(r * COUT), dx, dz, idesc, !(first && j == 0));
          }
        }
        tc_commit(&empty_bar[stage]);
      }
      first = false;
      __syncwarp();
      if (++stage == kWgStages) {
        stage = 0;
        phase ^= 1;
      }
    }
    if (lane == 0) tc_commit(done_bar);
    __syncwarp();
  } else {
    // drain: accumulator row m = s * CIN + ci (filter column s, input channel ci), column = co
    mbar_wait(done_bar, 0);
    tc_fence_after();
    if ((int)blockIdx.x < g.num_tiles) {
      const int quarter = warp & 3;
      const int m = quarter * 32 + lane;
      const int s = m / CIN, ci = m - s * CIN;
      constexpr int CW = COUT < 32 ? COUT : 32;
#pragma unroll 1
      for (int r = 0; r < 3; ++r) {
#pragma unroll 1
        for (int c = 0; c < COUT / CW; ++c) {
          uint32_t raw[32];
          const uint32_t ta = tmem_base + (uint32_t)(r * COUT + c * CW) + ((uint32_t)(quarter * 32) << 16);
          if (CW == 32) tc_ld32(ta, raw);
          else tc_ld16(ta, raw);
          tc_wait_ld();
          if (s < 3) {
#pragma unroll
            for (int j = 0; j < CW; ++j)
              atomicAdd(dw + ((long)((c * CW + j) * 3 + r) * 3 + s) * CIN + ci, __uint_as_float(raw[j]));
          }
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tc_dealloc(tmem_base, kCols);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// first layer (3 input channels) on CUDA cores
// ------------------------------------------------------------------------------------------------------------------
// y[n, gy, gx, 0..15] (bf16 NHWC) = bias + conv3x3(x fp32 NCHW [N,3,H,W]) ; statistics of the stored values
__global__ void __launch_bounds__(256)
sc_conv1_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias, bf16* __restrict__ y,
                int N, int H, int W, float* col_sum, float* col_sumsq) {
  __shared__ float xs[3][18][18];
  __shared__ __align__(16) float ws[27][16];
  __shared__ float bs[16];
  __shared__ float red[8][32];
  const int tiles_x = (W + 15) >> 4, tiles_y = (H + 15) >> 4;
  const long items = (long)N * tiles_x * tiles_y;
  const int lx = threadIdx.x & 15, ly = threadIdx.x >> 4;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 27 * 16; i += 256) {
    const int co = i & 15, k = i >> 4;                 // k = ci * 9 + tap ; torch layout w[co][ci][r][s]
    ws[k][co] = w[co * 27 + k];
  }
  if (threadIdx.x < 16) bs[threadIdx.x] = bias ? bias[threadIdx.x] : 0.f;
  float tot = 0.f;        // lane l < 16: sum of channel l ; lane l >= 16: sum of squares of channel l - 16
  for (long it = blockIdx.x; it < items; it += gridDim.x) {
    const int n = (int)(it / (tiles_x * tiles_y));
    const int tr = (int)(it - (long)n * tiles_x * tiles_y);
    const int tx0 = (tr % tiles_x) << 4, ty0 = (tr / tiles_x) << 4;
    __syncthreads();
    for (int idx = threadIdx.x; idx < 3 * 324; idx += 256) {
      const int ci = idx / 324, rem = idx - ci * 324;
      const int yy = rem / 18, xx = rem - yy * 18;
      const int gy = ty0 + yy - 1, gx = tx0 + xx - 1;
      xs[ci][yy][xx] = (gy >= 0 && gy < H && gx >= 0 && gx < W) ? x[(((long)n * 3 + ci) * H + gy) * W + gx] : 0.f;
    }
    __syncthreads();
    float acc[16];
#pragma unroll
    for (int o = 0; o < 16; ++o) acc[o] = bs[o];
#pragma unroll
    for (int ci = 0; ci < 3; ++ci) {
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const float v = xs[ci][ly + tap / 3][lx + tap % 3];
        const float4* wp = reinterpret_cast<const float4*>(&ws[ci * 9 + tap][0]);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 wv = wp[q];
          acc[q * 4 + 0] = fmaf(v, wv.x, acc[q * 4 + 0]);
          acc[q * 4 + 1] = fmaf(v, wv.y, acc[q * 4 + 1]);
          acc[q * 4 + 2] = fmaf(v, wv.z, acc[q * 4 + 2]);
          acc[q * 4 + 3] = fmaf(v, wv.w, acc[q * 4 + 3]);
        }
      }
    }
    const int gy = ty0 + ly, gx = tx0 + lx;
    const bool ok = gy < H && gx < W;
    uint32_t pk[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) pk[j] = pack_bf16x2(acc[2 * j], acc[2 * j + 1]);
    if (ok) {
      uint4* dp = reinterpret_cast<uint4*>(y + (((long)n * H + gy) * W + gx) * 16);
      dp[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      dp[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
    }
    if (col_sum != nullptr) {
      float v[32];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float lo = ok ? __uint_as_float(pk[j] << 16) : 0.f, hi = ok ? __uint_as_float(pk[j] & 0xffff0000u) : 0.f;
        v[2 * j] = lo;
        v[2 * j + 1] = hi;
        v[16 + 2 * j] = lo * lo;
        v[16 + 2 * j + 1] = hi * hi;
      }
      tot += colsum32(v, lane);
    }
  }
  if (col_sum != nullptr) {
    red[warp][lane] = tot;
    __syncthreads();
    if (threadIdx.x < 32) {
      float a = 0.f;
#pragma unroll
      for (int w8 = 0; w8 < 8; ++w8) a += red[w8][threadIdx.x];
      atomicAdd(threadIdx.x < 16 ? col_sum + threadIdx.x : col_sumsq + (threadIdx.x - 16), a);
    }
  }
}

// dw[co][ci][r][s] (fp32, torch layout, ACCUMULATED) += sum_pixels dz[n, y, x, co] (bf16 NHWC, 16 channels) * x[n, ci, y+r-1, x+s-1]
// thread (pixel group pg of 8, lane): lane -> 4 output channels x 4 (ci, tap) pairs = 16 register accumulators over the
// group's 32 pixels of every 16 x 16 tile the CTA walks
__global__ void __launch_bounds__(256)
sc_conv1_wgrad_kernel(const float* __restrict__ x, const bf16* __restrict__ dz, float* __restrict__ dw, int N, int H, int W) {
  __shared__ float xs[3][18][18];
  __shared__ __align__(16) float ds[256][16];
  const int tiles_x = (W + 15) >> 4, tiles_y = (H + 15) >> 4;
  const long items = (long)N * tiles_x * tiles_y;
  const int pg = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cob = lane & 3, jb = lane >> 2;               // channels 4 cob .. 4 cob + 3 ; pairs 4 jb .. 4 jb + 3 (of 27)
  const bool active = jb < 7;
  int xoff[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int j = min(jb * 4 + q, 26);
    const int ci = j / 9, tap = j - ci * 9;
    xoff[q] = (ci * 18 + tap / 3) * 18 + tap % 3;
  }
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
  const float* xsf = &xs[0][0][0];
  for (long it = blockIdx.x; it < items; it += gridDim.x) {
    const int n = (int)(it / (tiles_x * tiles_y));
    const int tr = (int)(it - (long)n * tiles_x * tiles_y);
    const int tx0 = (tr % tiles_x) << 4, ty0 = (tr / tiles_x) << 4;
    __syncthreads();
    for (int idx = threadIdx.x; idx < 3 * 324; idx += 256) {
      const int ci = idx / 324, rem = idx - ci * 324;
      const int yy = rem / 18, xx = rem - yy * 18;
      const int gy = ty0 + yy - 1, gx = tx0 + xx - 1;
      xs[ci][yy][xx] = (gy >= 0 && gy < H && gx >= 0 && gx < W) ? x[(((long)n * 3 + ci) * H + gy) * W + gx] : 0.f;
    }
    {
      const int px = threadIdx.x & 15, py = threadIdx.x >> 4;
      const int gy = ty0 + py, gx = tx0 + px;
      uint4 a = make_uint4(0, 0, 0, 0), b = a;
      if (gy < H && gx < W) {
        const uint4* p = reinterpret_cast<const uint4*>(dz + (((long)n * H + gy) * W + gx) * 16);
        a = __ldg(p);
        b = __ldg(p + 1);
      }
      const uint32_t u[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        ds[threadIdx.x][2 * j] = __uint_as_float(u[j] << 16);
        ds[threadIdx.x][2 * j + 1] = __uint_as_float(u[j] & 0xffff0000u);
      }
    }
    __syncthreads();
    if (active) {
#pragma unroll 4
      for (int i = 0; i < 32; ++i) {
        const int p = pg * 32 + i;                       // pixel (py = p >> 4, px = p & 15)
        const float4 d4 = *reinterpret_cast<const float4*>(&ds[p][cob * 4]);
        const int base = (p >> 4) * 18 + (p & 15);
        float xv[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) xv[q] = xsf[base + xoff[q]];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          acc[0][q] = fmaf(d4.x, xv[q], acc[0][q]);
          acc[1][q] = fmaf(d4.y, xv[q], acc[1][q]);
          acc[2][q] = fmaf(d4.z, xv[q], acc[2][q]);
          acc[3][q] = fmaf(d4.w, xv[q], acc[3][q]);
        }
      }
    }
  }
  if (active) {
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int j = jb * 4 + q;
        if (j < 27) atomicAdd(dw + (cob * 4 + a) * 27 + j, acc[a][q]);
      }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// BatchNorm finalisation, BN + ReLU (+ pool) forward / backward, layout changes
// ------------------------------------------------------------------------------------------------------------------
// The raw conv outputs are stored WITHOUT the conv bias (train-mode BatchNorm subtracts it again; adding it in the conv epilogue
// costs 32 loads per TMEM chunk): the bias only shifts the mean, so it is folded in here -- running_mean tracks mean(raw) + bias,
// and with running statistics (eval) the effective mean of the bias-free tensor is running_mean - bias.
__global__ void sc_bn_finalize_kernel(const float* sum, const float* sumsq, const float* conv_bias, const float* gamma,
                                      const float* beta, float* running_mean, float* running_var, float inv_count, float unbias,
                                      float eps, float momentum, int train, float* scale, float* shift, float* mean_out,
                                      float* rstd_out, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float b = conv_bias != nullptr ? conv_bias[c] : 0.f;
  float mean, var;
  if (train) {
    mean = sum[c] * inv_count;
    var = fmaxf(sumsq[c] * inv_count - mean * mean, 0.f);
    if (running_mean != nullptr) {
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (mean + b);
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * var * unbias;
    }
  } else {
    mean = running_mean[c] - b;
    var = running_var[c];
  }
  const float rstd = rsqrtf(var + eps);
  const float sc = gamma[c] * rstd;
  scale[c] = sc;
  shift[c] = beta[c] - mean * sc;
  mean_out[c] = mean;
  rstd_out[c] = rstd;
}

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    f[2 * j] = __uint_as_float(w[j] << 16);
    f[2 * j + 1] = __uint_as_float(w[j] & 0xffff0000u);
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
}
__device__ __forceinline__ void load8f(const float* p, float (&f)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

// y[n, yo, xo, c] = max over the POOL x POOL window of relu(raw * scale[c] + shift[c]) ; thread = (output pixel, 8 channels)
template <int POOL>
__global__ void __launch_bounds__(256)
sc_act_pool_fwd_kernel(const bf16* __restrict__ raw, const float* __restrict__ scale, const float* __restrict__ shift,
                       bf16* __restrict__ y, int N, int H, int W, int C) {
  const int cgs = C >> 3;
  const int Ho = H / POOL, Wo = W / POOL;
  const long total = (long)N * Ho * Wo * cgs;
  const long stride = (long)gridDim.x * blockDim.x;          // a multiple of 8 -> a thread's channel group is fixed
  long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int cg = (int)(idx % cgs);
  float sc[8], sh[8];
  load8f(scale + cg * 8, sc);
  load8f(shift + cg * 8, sh);
  const int sh_cg = cgs == 8 ? 3 : (cgs == 4 ? 2 : 1);     // 32-bit index math: 64-bit divisions made these passes ALU bound
  for (; idx < total; idx += stride) {
    const uint32_t pix = (uint32_t)(idx >> sh_cg);
    const uint32_t t = pix / (uint32_t)Wo;
    const uint32_t xo = pix - t * (uint32_t)Wo;
    const uint32_t n = t / (uint32_t)Ho;
    const uint32_t yo = t - n * (uint32_t)Ho;
    const bf16* win = raw + (((long)n * H + yo * POOL) * W + xo * POOL) * C + cg * 8;
    float best[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) best[j] = 0.f;                // relu floor
#pragma unroll
    for (int dy = 0; dy < POOL; ++dy)
#pragma unroll
      for (int dx = 0; dx < POOL; ++dx) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(win + ((long)dy * W + dx) * C));
        float f[8];
        unpack8(u, f);
#pragma unroll
        for (int j = 0; j < 8; ++j) best[j] = fmaxf(best[j], fmaf(f[j], sc[j], sh[j]));
      }
    *reinterpret_cast<uint4*>(y + (long)pix * C + cg * 8) = pack8(best);
  }
}

// The gradient arriving at relu(bn(raw)) after un-pooling: window position k gets dy when it is the FIRST maximum of the
// window (torch's max_pool2d backward) and its activation is positive, else 0.  One kernel body serves both passes of the
// train-mode BatchNorm backward:
//   APPLY = false   s1[c] += sum dpre ; s2[c] += sum dpre * xhat            (dbeta, dgamma; ACCUMULATED)
//   APPLY = true    dz = scale * (dpre - s1/M - xhat * s2/M) (train) | scale * dpre (eval), full resolution, every position
// A thread owns 8 channels; all 128-bit loads of its pixel(s) (POOL = 1: two pixels, 4 loads; POOL = 2: one pooled pixel,
// 5 loads) are issued before any math (the first version, one pixel in flight at 100+ registers, ran at a third of the copy bandwidth: latency bound); the
// window stays packed in registers and is unpacked once to find the arg-max and once to emit.
template <int POOL, bool APPLY>
__global__ void __launch_bounds__(256, 2)
sc_act_pool_bwd_kernel(const bf16* __restrict__ raw, const bf16* __restrict__ dyp, const float* __restrict__ scale,
                       const float* __restrict__ shift, const float* __restrict__ mean, const float* __restrict__ rstd,
                       float* __restrict__ s1, float* __restrict__ s2, float inv_count, int train, bf16* __restrict__ dz, int N,
                       int H, int W, int C) {
  constexpr int KW = POOL * POOL;
  __shared__ float red[2][64];
  const int cgs = C >> 3;
  const int Ho = H / POOL, Wo = W / POOL;
  const long total = (long)N * Ho * Wo * cgs;
  const long stride = (long)gridDim.x * blockDim.x;        // a multiple of 8 -> a thread's channel group is fixed
  long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int cg = (int)(idx % cgs);
  float sc[8], sh[8], x1[8], x0[8], m1[8], m2[8];           // xhat = raw * x1 + x0
  load8f(scale + cg * 8, sc);
  load8f(shift + cg * 8, sh);
  load8f(rstd + cg * 8, x1);
  load8f(mean + cg * 8, x0);
#pragma unroll
  for (int j = 0; j < 8; ++j) x0[j] = -x0[j] * x1[j];
  if (APPLY) {
    load8f(s1 + cg * 8, m1);
    load8f(s2 + cg * 8, m2);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      m1[j] = train ? m1[j] * inv_count : 0.f;
      m2[j] = train ? m2[j] * inv_count : 0.f;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) m1[j] = m2[j] = 0.f;       // accumulators
    if (threadIdx.x < 128) red[threadIdx.x >> 6][threadIdx.x & 63] = 0.f;
    __syncthreads();
  }
  const int sh_cg = cgs == 8 ? 3 : (cgs == 4 ? 2 : 1);
  auto win_base = [&](uint32_t pix) {        // element offset of window position 0 (32-bit index math, one division pair per pixel)
    if (POOL == 1) return (long)pix * C + cg * 8;
    const uint32_t t = pix / (uint32_t)Wo;
    const uint32_t xo = pix - t * (uint32_t)Wo;
    const uint32_t n = t / (uint32_t)Ho;
    const uint32_t yo = t - n * (uint32_t)Ho;
    return (((long)n * H + yo * POOL) * W + xo * POOL) * C + cg * 8;
  };
  auto win_off = [&](int k) { return ((long)(k / POOL) * W + (k % POOL)) * C; };
  auto process = [&](long base, const uint4 (&rw)[KW], const uint4& dyu) {
    float dy[8], best[8];
    int arg[8];
    unpack8(dyu, dy);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      best[j] = -1.f;
      arg[j] = 0;
    }
#pragma unroll
    for (int k = 0; k < KW; ++k) {
      float f[8];
      unpack8(rw[k], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float a = fmaxf(fmaf(f[j], sc[j], sh[j]), 0.f);
        if (a > best[j]) {
          best[j] = a;
          arg[j] = k;
        }
      }
    }
#pragma unroll
    for (int k = 0; k < KW; ++k) {
      float f[8], o[8];
      unpack8(rw[k], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float dpre = (arg[j] == k && best[j] > 0.f) ? dy[j] : 0.f;
        const float xh = fmaf(f[j], x1[j], x0[j]);
        if (APPLY) {
          o[j] = sc[j] * (dpre - m1[j] - xh * m2[j]);
        } else {
          m1[j] += dpre;
          m2[j] = fmaf(dpre, xh, m2[j]);
        }
      }
      if (APPLY) *reinterpret_cast<uint4*>(dz + base + win_off(k)) = pack8(o);
    }
  };
  if constexpr (POOL == 1) {           // two pixels in flight (4 loads)
    for (; idx < total; idx += 2 * stride) {
      const long ib = idx + stride;
      const bool has_b = ib < total;
      const uint32_t pa = (uint32_t)(idx >> sh_cg), pb = has_b ? (uint32_t)(ib >> sh_cg) : pa;
      const long ba = win_base(pa), bb = win_base(pb);
      uint4 ra[KW], rb[KW];
      ra[0] = __ldg(reinterpret_cast<const uint4*>(raw + ba));
      const uint4 da = __ldg(reinterpret_cast<const uint4*>(dyp + ba));
      rb[0] = __ldg(reinterpret_cast<const uint4*>(raw + bb));
      const uint4 db = __ldg(reinterpret_cast<const uint4*>(dyp + bb));
      process(ba, ra, da);
      if (has_b) process(bb, rb, db);
    }
  } else {                             // one pooled pixel = 5 loads in flight
    for (; idx < total; idx += stride) {
      const uint32_t pa = (uint32_t)(idx >> sh_cg);
      const long ba = win_base(pa);
      uint4 ra[KW];
#pragma unroll
      for (int k = 0; k < KW; ++k) ra[k] = __ldg(reinterpret_cast<const uint4*>(raw + ba + win_off(k)));
      const uint4 da = __ldg(reinterpret_cast<const uint4*>(dyp + (long)pa * C + cg * 8));
      process(ba, ra, da);
    }
  }
  if (!APPLY) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      atomicAdd(&red[0][cg * 8 + j], m1[j]);
      atomicAdd(&red[1][cg * 8 + j], m2[j]);
    }
    __syncthreads();
    if (threadIdx.x < C) {
      atomicAdd(s1 + threadIdx.x, red[0][threadIdx.x]);
      atomicAdd(s2 + threadIdx.x, red[1][threadIdx.x]);
    }
  }
}

// x fp32 NCHW [N,3,H,W] -> bf16 NHWC [N,H,W,16] with channels 3..15 zero: the first layer then runs on the same tensor-core
// kernels as the others (the 0..255 / 0..1 pixel values lose nothing that the bf16 conv input of autocast would keep)
__global__ void __launch_bounds__(256)
sc_pack_input_kernel(const float* __restrict__ x, bf16* __restrict__ y, long pixels_per_img, long total) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long n = i / pixels_per_img, p = i - n * pixels_per_img;
    const float* src = x + n * 3 * pixels_per_img + p;
    const uint32_t w0 = pack_bf16x2(src[0], src[pixels_per_img]);
    const uint32_t w1 = pack_bf16x2(src[2 * pixels_per_img], 0.f);
    uint4* dst = reinterpret_cast<uint4*>(y + i * 16);
    dst[0] = make_uint4(w0, w1, 0u, 0u);
    dst[1] = make_uint4(0u, 0u, 0u, 0u);
  }
}

__device__ __forceinline__ uint32_t sc_mix32(uint64_t z) {          // same generator as tail_ops.cu::dropout_kernel
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return (uint32_t)((z ^ (z >> 31)) >> 32);
}

// feat[n][c * HW + p] (bf16) = drop(act[n][p][c])  -- the channel-major flatten of nb:186 + nn.Dropout (mask = hash of the
// flat output index, replayed by the backward).  CTA = one image; a 32 x 33 shared-memory tile transposes.
__global__ void __launch_bounds__(256)
sc_nhwc_to_chw_kernel(const bf16* __restrict__ act, bf16* __restrict__ feat, int HW, int C, float p_drop,
                      unsigned long long seed,
                      const unsigned long long* __restrict__ seed_offset) {
  if (seed_offset != nullptr) seed += *seed_offset;
  __shared__ float tile[32][33];
  const long n = blockIdx.y;
  const int p0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const float inv_keep = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
  const uint32_t thresh = (uint32_t)((double)p_drop * 4294967296.0);
  for (int c0 = 0; c0 < C; c0 += 32) {
    for (int i = ty; i < 32; i += 8) {
      const int p = p0 + i, c = c0 + tx;
      tile[i][tx] = (p < HW && c < C) ? __bfloat162float(act[(n * HW + p) * C + c]) : 0.f;
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
      const int c = c0 + i, p = p0 + tx;
      if (c < C && p < HW) {
        const long o = (n * C + c) * HW + p;
        float v = tile[tx][i];
        if (p_drop > 0.f) v = sc_mix32(seed * 0xD1342543DE82EF95ull + (uint64_t)o) >= thresh ? v * inv_keep : 0.f;
        feat[o] = __float2bfloat16_rn(v);
      }
    }
    __syncthreads();
  }
}

// dact[n][p][c] (bf16) = drop'(dfeat[n][c * HW + p]) (fp32 or bf16 in): inverse layout change + the same dropout mask
__device__ __forceinline__ float ld_as_float(const float* p) { return *p; }
__device__ __forceinline__ float ld_as_float(const bf16* p) { return __bfloat162float(*p); }
template <typename InT>
__global__ void __launch_bounds__(256)
sc_chw_to_nhwc_kernel(const InT* __restrict__ dfeat, bf16* __restrict__ dact, int HW, int C, float p_drop,
                      unsigned long long seed,
                      const unsigned long long* __restrict__ seed_offset) {
  if (seed_offset != nullptr) seed += *seed_offset;
  __shared__ float tile[32][33];
  const long n = blockIdx.y;
  const int p0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const float inv_keep = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
  const uint32_t thresh = (uint32_t)((double)p_drop * 4294967296.0);
  for (int c0 = 0; c0 < C; c0 += 32) {
    for (int i = ty; i < 32; i += 8) {
      const int c = c0 + i, p = p0 + tx;
      float v = 0.f;
      if (c < C && p < HW) {
        const long o = (n * C + c) * HW + p;
        v = ld_as_float(dfeat + o);
        if (p_drop > 0.f) v = sc_mix32(seed * 0xD1342543DE82EF95ull + (uint64_t)o) >= thresh ? v * inv_keep : 0.f;
      }
      tile[i][tx] = v;
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
      const int p = p0 + i, c = c0 + tx;
      if (p < HW && c < C) dact[(n * HW + p) * C + c] = __float2bfloat16_rn(tile[tx][i]);
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
std::once_flag g_once;

int load_encode() {
  std::call_once(g_once, [] {
    cudaDriverEntryPointQueryResult qres;
    void* fn = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  });
  if (g_encode == nullptr) {
    b2_set_error("cuTensorMapEncodeTiled driver entry point unavailable (no CUDA driver / GPU?)");
    return -2;
  }
  return 0;
}

CUtensorMapSwizzle swizzle_of(int row_bytes) {
  return row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}

// NHWC activation [N, H, W, C] -> box {C, box_w, box_h, 1}
int make_act_map(CUtensorMap* m, const void* p, int N, int H, int W, int C, int box_w, int box_h, const char* who) {
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)C, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult cr = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(p), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_of(C * 2), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) {
    b2_set_error("%s: cuTensorMapEncodeTiled(activation) failed (%d): N=%d H=%d W=%d C=%d box %dx%d", who, (int)cr, N, H, W, C,
                 box_w, box_h);
    return -3;
  }
  return 0;
}

// Picks the rows per tile: most useful rows per 128-row accumulator block among the candidates that fit TMEM and shared memory
int make_sc_geom(ScGeom* g, int N, int H, int W, int Cin, int Cout) {
  g->N = N; g->H = H; g->W = W;
  g->Wp = W + 2;
  if (g->Wp > 256) return -1;
  const int rb = Cin * 2;
  const int nb_max = 256 / Cout < 16 ? 256 / Cout : 16;          // two accumulator sets of NB x Cout columns in 512
  const int w_bytes = kTaps * ((Cout * rb + 1023) / 1024 * 1024) + 8 * 16 * 33 * 4;
  double best = -1.0;
  int best_th = 0;
  for (int th = 1; th <= H && th + 2 <= 256; ++th) {
    const int nb = (th * g->Wp + 127) / 128;
    if (nb > nb_max) break;
    const int reach = nb * 128 + 2 * g->Wp + 2;
    const int halo = (th + 2) * g->Wp;
    const int stage = ((reach > halo ? reach : halo) * rb + 1023) / 1024 * 1024;
    if (w_bytes + 2 * stage + 4096 > 220 * 1024) break;
    const int tiles = (H + th - 1) / th;
    const double eff = (double)H * W / ((double)tiles * nb * 128);
    if (eff > best + 1e-9) {
      best = eff;
      best_th = th;
    }
  }
  if (best_th == 0) return -1;
  g->TH = best_th;
  g->NB = (g->TH * g->Wp + 127) / 128;
  g->Mt = g->TH * g->Wp;
  g->tiles_per_img = (H + g->TH - 1) / g->TH;
  const long tiles = (long)N * g->tiles_per_img;
  if (tiles >= (1L << 31) || (unsigned long)tiles * g->tiles_per_img >= (1ul << 32)) return -1;
  g->num_tiles = (int)tiles;
  g->halo_px = (g->TH + 2) * g->Wp;
  const int reach = g->NB * 128 + 2 * g->Wp + 2;
  g->stage_bytes = ((reach > g->halo_px ? reach : g->halo_px) * rb + 1023) / 1024 * 1024;
  int stages = (220 * 1024 - w_bytes - 4096) / g->stage_bytes;
  g->stages = stages > 4 ? 4 : stages;
  g->magic_wp = (uint32_t)(((1ull << 32) + g->Wp - 1) / g->Wp);
  g->magic_tpi = g->tiles_per_img == 1 ? 0u : (uint32_t)(((1ull << 32) + g->tiles_per_img - 1) / g->tiles_per_img);
  return g->stages >= 2 ? 0 : -1;
}

template <int CIN, int COUT>
int launch_sc_conv(const void* x, int N, int H, int W, const void* w, void* y, const float* bias, float* col_sum,
                   float* col_sumsq, cudaStream_t st, const char* who) {
  ScGeom g;
  B2_ARG_CHECK(make_sc_geom(&g, N, H, W, CIN, COUT) == 0, "%s: unsupported shape N=%d H=%d W=%d", who, N, H, W);
  CUtensorMap tx, tw;
  if (int r = make_act_map(&tx, x, N, H, W, CIN, g.Wp, g.TH + 2, who)) return r;
  {
    cuuint64_t dims[2] = {(cuuint64_t)(kTaps * CIN), (cuuint64_t)COUT};
    cuuint64_t strides[1] = {(cuuint64_t)(kTaps * CIN) * 2};
    cuuint32_t box[2] = {(cuuint32_t)CIN, (cuuint32_t)COUT};
    cuuint32_t estr[2] = {1, 1};
    CUresult cr = g_encode(&tw, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_of(CIN * 2), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) {
      b2_set_error("%s: cuTensorMapEncodeTiled(w) failed (%d)", who, (int)cr);
      return -3;
    }
  }
  const int w_slot = (COUT * CIN * 2 + 1023) / 1024 * 1024;
  const int smem = kTaps * w_slot + g.stages * g.stage_bytes + 128 + 8 * 2 * COUT * 4 + 8 * 16 * 33 * 4 + 1024;
  static B2PerDeviceMax attr;
  if (attr.below(smem)) {
    B2_CUDA_CHECK(cudaFuncSetAttribute(sc_conv_kernel<CIN, COUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr.set(smem);
  }
  const int grid = g.num_tiles < b2_num_sms() ? g.num_tiles : b2_num_sms();
  sc_conv_kernel<CIN, COUT><<<grid, kConvThreads, smem, st>>>(tx, tw, (bf16*)y, g, bias, col_sum, col_sumsq);
  B2_LAUNCH_CHECK("sc_conv_kernel");
  return 0;
}

template <int CIN, int COUT>
int launch_sc_wgrad(const void* x, const void* dz, int N, int H, int W, float* dw, cudaStream_t st, const char* who) {
  WgGeomSc g;
  g.N = N; g.H = H; g.W = W;
  g.Wp = W + 2;
  B2_ARG_CHECK(g.Wp <= 256, "%s: W too large", who);
  // rows per tile: as many as three stages of (x halo + dz) allow, at most 254 halo rows per TMA box
  int th = H < 254 ? H : 254;
  for (;; --th) {
    B2_ARG_CHECK(th >= 1, "%s: image rows do not fit shared memory (W=%d)", who, W);
    const int mt = th * g.Wp, ks = (mt + 15) / 16;
    const int xb = ((ks * 16 + 2 * g.Wp + 8) * CIN * 2 + 1023) / 1024 * 1024;
    const int zb = (ks * 16 * COUT * 2 + 1023) / 1024 * 1024;
    if (kWgStages * (xb + zb) + 2048 <= 220 * 1024) {
      g.TH = th; g.Mt = mt; g.ksteps = ks; g.x_bytes = xb; g.z_bytes = zb;
      break;
    }
  }
  g.tiles_per_img = (H + g.TH - 1) / g.TH;
  const long tiles = (long)N * g.tiles_per_img;
  B2_ARG_CHECK(tiles < (1L << 31) && (unsigned long)tiles * g.tiles_per_img < (1ul << 32), "%s: too many tiles", who);
  g.num_tiles = (int)tiles;
  g.halo_px = (g.TH + 2) * g.Wp;
  g.magic_tpi = g.tiles_per_img == 1 ? 0u : (uint32_t)(((1ull << 32) + g.tiles_per_img - 1) / g.tiles_per_img);
  CUtensorMap tx, tz;
  if (int r = make_act_map(&tx, x, N, H, W, CIN, g.Wp, g.TH + 2, who)) return r;
  if (int r = make_act_map(&tz, dz, N, H, W, COUT, g.Wp, g.TH, who)) return r;
  const int smem = kWgStages * (g.x_bytes + g.z_bytes) + 1024 + 1024;
  static B2PerDeviceMax attr;
  if (attr.below(smem)) {
    B2_CUDA_CHECK(cudaFuncSetAttribute(sc_wgrad_kernel<CIN, COUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr.set(smem);
  }
  const int grid = g.num_tiles < b2_num_sms() ? g.num_tiles : b2_num_sms();
  sc_wgrad_kernel<CIN, COUT><<<grid, kWgThreads, smem, st>>>(tx, tz, g, dw);
  B2_LAUNCH_CHECK("sc_wgrad_kernel");
  return 0;
}

int ew_grid_sc(long items) {
  const long cap = (long)b2_num_sms() * 8;
  long blocks = (items + 255) / 256;
  return (int)(blocks < cap ? (blocks < 1 ? 1 : blocks) : cap);
}

}  // namespace

// ---- C ABI (include/b200lrcn.h, "small-CNN tensor-core path") ----------------------------------------------------------

// y [N,H,W,Cout] bf16 = conv3x3(x [N,H,W,Cin] bf16, w [Cout][3][3][Cin] bf16) (+ bias), stride 1, pad 1; optional per-channel
// sum / sum of squares of the stored values (ACCUMULATED).  (Cin, Cout) in {(16,32), (32,64), (64,32), (32,16)}: the forward
// convs of the notebook CNN and their data gradients (w = the flipped, transposed filter)
B2_API int b2_sc_conv3x3_bf16(const void* x, int N, int H, int W, int Cin, const void* w, int Cout, void* y, const float* bias,
                              float* col_sum, float* col_sumsq, void* stream) {
  const char* who = "b2_sc_conv3x3_bf16";
  B2_ARG_CHECK(x && w && y && N > 0 && H > 0 && W > 0, "%s: null pointer or empty shape", who);
  B2_ARG_CHECK((col_sum == nullptr) == (col_sumsq == nullptr), "%s: col_sum and col_sumsq go together", who);
  B2_ARG_CHECK(((uintptr_t)x & 15) == 0 && ((uintptr_t)w & 15) == 0 && ((uintptr_t)y & 31) == 0, "%s: alignment", who);
  if (int r = load_encode()) return r;
  cudaStream_t st = (cudaStream_t)stream;
  if (Cin == 16 && Cout == 16) return launch_sc_conv<16, 16>(x, N, H, W, w, y, bias, col_sum, col_sumsq, st, who);
  if (Cin == 16 && Cout == 32) return launch_sc_conv<16, 32>(x, N, H, W, w, y, bias, col_sum, col_sumsq, st, who);
  if (Cin == 32 && Cout == 64) return launch_sc_conv<32, 64>(x, N, H, W, w, y, bias, col_sum, col_sumsq, st, who);
  if (Cin == 64 && Cout == 32) return launch_sc_conv<64, 32>(x, N, H, W, w, y, bias, col_sum, col_sumsq, st, who);
  if (Cin == 32 && Cout == 16) return launch_sc_conv<32, 16>(x, N, H, W, w, y, bias, col_sum, col_sumsq, st, who);
  b2_set_error("%s: unsupported channel pair (%d -> %d)", who, Cin, Cout);
  return -1;
}

// dw [Cout][3][3][Cin] fp32 (ACCUMULATED, caller zeroes) += sum_pixels dz [N,H,W,Cout] x shifted x [N,H,W,Cin]; (16,32) or (32,64)
B2_API int b2_sc_conv3x3_wgrad_bf16(const void* x, const void* dz, int N, int H, int W, int Cin, int Cout, float* dw,
                                    void* stream) {
  const char* who = "b2_sc_conv3x3_wgrad_bf16";
  B2_ARG_CHECK(x && dz && dw && N > 0 && H > 0 && W > 0, "%s: null pointer or empty shape", who);
  B2_ARG_CHECK(((uintptr_t)x & 15) == 0 && ((uintptr_t)dz & 15) == 0, "%s: alignment", who);
  if (int r = load_encode()) return r;
  cudaStream_t st = (cudaStream_t)stream;
  if (Cin == 16 && Cout == 16) return launch_sc_wgrad<16, 16>(x, dz, N, H, W, dw, st, who);
  if (Cin == 16 && Cout == 32) return launch_sc_wgrad<16, 32>(x, dz, N, H, W, dw, st, who);
  if (Cin == 32 && Cout == 64) return launch_sc_wgrad<32, 64>(x, dz, N, H, W, dw, st, who);
  b2_set_error("%s: unsupported channel pair (%d -> %d)", who, Cin, Cout);
  return -1;
}

// first layer: x fp32 NCHW [N,3,H,W], w fp32 [16][3][3][3] (torch layout) -> y bf16 NHWC [N,H,W,16] (+ statistics, ACCUMULATED)
B2_API int b2_sc_conv1_fwd(const float* x, const float* w, const float* bias, void* y, int N, int H, int W, float* col_sum,
                           float* col_sumsq, void* stream) {
  B2_ARG_CHECK(x && w && y && N > 0 && H > 0 && W > 0, "b2_sc_conv1_fwd: null pointer or empty shape");
  B2_ARG_CHECK((col_sum == nullptr) == (col_sumsq == nullptr), "b2_sc_conv1_fwd: col_sum and col_sumsq go together");
  const long items = (long)N * ((W + 15) / 16) * ((H + 15) / 16);
  const long cap = (long)b2_num_sms() * 4;
  sc_conv1_kernel<<<(unsigned)(items < cap ? items : cap), 256, 0, (cudaStream_t)stream>>>(x, w, bias, (bf16*)y, N, H, W, col_sum,
                                                                                           col_sumsq);
  B2_LAUNCH_CHECK("sc_conv1_kernel");
  return 0;
}

// dw fp32 [16][3][3][3] (ACCUMULATED) from x fp32 NCHW and dz bf16 NHWC [N,H,W,16]
B2_API int b2_sc_conv1_wgrad(const float* x, const void* dz, float* dw, int N, int H, int W, void* stream) {
  B2_ARG_CHECK(x && dz && dw && N > 0 && H > 0 && W > 0, "b2_sc_conv1_wgrad: null pointer or empty shape");
  const long items = (long)N * ((W + 15) / 16) * ((H + 15) / 16);
  const long cap = (long)b2_num_sms() * 4;
  sc_conv1_wgrad_kernel<<<(unsigned)(items < cap ? items : cap), 256, 0, (cudaStream_t)stream>>>(x, (const bf16*)dz, dw, N, H, W);
  B2_LAUNCH_CHECK("sc_conv1_wgrad_kernel");
  return 0;
}

// x fp32 NCHW [N,3,H,W] -> bf16 NHWC [N,H,W,16] (channels 3..15 zero)
B2_API int b2_sc_pack_input(const float* x, void* y, int N, int H, int W, void* stream) {
  B2_ARG_CHECK(x && y && N > 0 && H > 0 && W > 0 && ((uintptr_t)y & 15) == 0, "b2_sc_pack_input: null pointer, empty shape or alignment");
  const long total = (long)N * H * W;
  sc_pack_input_kernel<<<ew_grid_sc(total), 256, 0, (cudaStream_t)stream>>>(x, (bf16*)y, (long)H * W, total);
  B2_LAUNCH_CHECK("sc_pack_input_kernel");
  return 0;
}

B2_API int b2_sc_bn_finalize(const float* sum, const float* sumsq, const float* conv_bias, const float* gamma, const float* beta,
                             float* running_mean, float* running_var, long count, float eps, float momentum, int train,
                             float* scale, float* shift, float* mean, float* rstd, int C, void* stream) {
  B2_ARG_CHECK(gamma && beta && scale && shift && mean && rstd && C > 0 && count > 0, "b2_sc_bn_finalize: null pointer or empty");
  B2_ARG_CHECK(train ? (sum && sumsq) : (running_mean && running_var), "b2_sc_bn_finalize: missing statistics");
  const float inv = (float)(1.0 / (double)count);
  const float unbias = count > 1 ? (float)((double)count / (double)(count - 1)) : 1.f;
  sc_bn_finalize_kernel<<<(C + 63) / 64, 64, 0, (cudaStream_t)stream>>>(sum, sumsq, conv_bias, gamma, beta, running_mean, running_var,
                                                                        inv, unbias, eps, momentum, train, scale, shift, mean, rstd,
                                                                        C);
  B2_LAUNCH_CHECK("sc_bn_finalize_kernel");
  return 0;
}

// y [N,H/pool,W/pool,C] = maxpool_pool(relu(raw * scale + shift)) ; NHWC bf16, C in {16, 32, 64}, pool in {1, 2}
B2_API int b2_sc_act_pool_fwd(const void* raw, const float* scale, const float* shift, void* y, int N, int H, int W, int C,
                              int pool, void* stream) {
  B2_ARG_CHECK(raw && scale && shift && y && N > 0 && H > 0 && W > 0, "b2_sc_act_pool_fwd: null pointer or empty shape");
  B2_ARG_CHECK((C == 16 || C == 32 || C == 64) && (pool == 1 || (pool == 2 && H % 2 == 0 && W % 2 == 0)),
               "b2_sc_act_pool_fwd: C in {16,32,64}, pool 1 or 2 (even H, W)");
  const long items = (long)N * (H / pool) * (W / pool) * (C / 8);
  B2_ARG_CHECK((long)N * H * W < (1L << 31), "b2_sc_act_pool_fwd: too many pixels for 32-bit index math");
  cudaStream_t st = (cudaStream_t)stream;
  if (pool == 2) sc_act_pool_fwd_kernel<2><<<ew_grid_sc(items), 256, 0, st>>>((const bf16*)raw, scale, shift, (bf16*)y, N, H, W, C);
  else sc_act_pool_fwd_kernel<1><<<ew_grid_sc(items), 256, 0, st>>>((const bf16*)raw, scale, shift, (bf16*)y, N, H, W, C);
  B2_LAUNCH_CHECK("sc_act_pool_fwd_kernel");
  return 0;
}

// two-pass backward of the same block: reduce -> s1 = dbeta, s2 = dgamma (ACCUMULATED), then dz (full resolution)
B2_API int b2_sc_act_pool_bwd_reduce(const void* raw, const void* dy, const float* scale, const float* shift, const float* mean,
                                     const float* rstd, float* s1, float* s2, int N, int H, int W, int C, int pool, void* stream) {
  B2_ARG_CHECK(raw && dy && scale && shift && mean && rstd && s1 && s2, "b2_sc_act_pool_bwd_reduce: null pointer");
  B2_ARG_CHECK((C == 16 || C == 32 || C == 64) && (pool == 1 || (pool == 2 && H % 2 == 0 && W % 2 == 0)),
               "b2_sc_act_pool_bwd_reduce: C in {16,32,64}, pool 1 or 2 (even H, W)");
  B2_ARG_CHECK((long)N * H * W < (1L << 31), "b2_sc_act_pool_bwd_reduce: too many pixels for 32-bit index math");
  const long items = (long)N * (H / pool) * (W / pool) * (C / 8);
  cudaStream_t st = (cudaStream_t)stream;
  if (pool == 2)
    sc_act_pool_bwd_kernel<2, false><<<ew_grid_sc(items), 256, 0, st>>>((const bf16*)raw, (const bf16*)dy, scale, shift, mean, rstd,
                                                                           s1, s2, 0.f, 1, nullptr, N, H, W, C);
  else
    sc_act_pool_bwd_kernel<1, false><<<ew_grid_sc(items / 2), 256, 0, st>>>((const bf16*)raw, (const bf16*)dy, scale, shift, mean, rstd,
                                                                           s1, s2, 0.f, 1, nullptr, N, H, W, C);
  B2_LAUNCH_CHECK("sc_act_pool_bwd_kernel<reduce>");
  return 0;
}

B2_API int b2_sc_act_pool_bwd_apply(const void* raw, const void* dy, const float* scale, const float* shift, const float* mean,
                                    const float* rstd, float* s1, float* s2, int train, void* dz, int N, int H, int W,
                                    int C, int pool, void* stream) {
  B2_ARG_CHECK(raw && dy && scale && shift && mean && rstd && s1 && s2 && dz, "b2_sc_act_pool_bwd_apply: null pointer");
  B2_ARG_CHECK((C == 16 || C == 32 || C == 64) && (pool == 1 || (pool == 2 && H % 2 == 0 && W % 2 == 0)),
               "b2_sc_act_pool_bwd_apply: C in {16,32,64}, pool 1 or 2 (even H, W)");
  B2_ARG_CHECK((long)N * H * W < (1L << 31), "b2_sc_act_pool_bwd_apply: too many pixels for 32-bit index math");
  const long items = (long)N * (H / pool) * (W / pool) * (C / 8);
  const float inv = (float)(1.0 / ((double)N * H * W));
  cudaStream_t st = (cudaStream_t)stream;
  if (pool == 2)
    sc_act_pool_bwd_kernel<2, true><<<ew_grid_sc(items), 256, 0, st>>>((const bf16*)raw, (const bf16*)dy, scale, shift, mean, rstd, s1,
                                                                          s2, inv, train, (bf16*)dz, N, H, W, C);
  else
    sc_act_pool_bwd_kernel<1, true><<<ew_grid_sc(items / 2), 256, 0, st>>>((const bf16*)raw, (const bf16*)dy, scale, shift, mean, rstd, s1,
                                                                          s2, inv, train, (bf16*)dz, N, H, W, C);
  B2_LAUNCH_CHECK("sc_act_pool_bwd_kernel<apply>");
  return 0;
}

// feat [N][C*HW] bf16 (channel-major flatten, nb:186) = dropout(act [N][HW][C] bf16) ; p = 0: plain layout change
B2_API int b2_sc_nhwc_to_chw(const void* act, void* feat, int N, int HW, int C, float p_drop, unsigned long long seed,
                             const unsigned long long* seed_offset, void* stream) {
  B2_ARG_CHECK(act && feat && N > 0 && HW > 0 && C > 0 && N <= 65535, "b2_sc_nhwc_to_chw: null pointer or bad shape");
  B2_ARG_CHECK(p_drop >= 0.f && p_drop < 1.f, "b2_sc_nhwc_to_chw: p must be in [0,1)");
  sc_nhwc_to_chw_kernel<<<dim3((HW + 31) / 32, N), 256, 0, (cudaStream_t)stream>>>((const bf16*)act, (bf16*)feat, HW, C, p_drop, seed, seed_offset);
  B2_LAUNCH_CHECK("sc_nhwc_to_chw_kernel");
  return 0;
}

// dact [N][HW][C] bf16 = dropout'(dfeat [N][C*HW] fp32 or bf16) with the mask of the forward call (same seed)
B2_API int b2_sc_chw_to_nhwc(const void* dfeat, int in_bf16, void* dact, int N, int HW, int C, float p_drop,
                             unsigned long long seed, const unsigned long long* seed_offset, void* stream) {
  B2_ARG_CHECK(dfeat && dact && N > 0 && HW > 0 && C > 0 && N <= 65535, "b2_sc_chw_to_nhwc: null pointer or bad shape");
  B2_ARG_CHECK(p_drop >= 0.f && p_drop < 1.f, "b2_sc_chw_to_nhwc: p must be in [0,1)");
  if (in_bf16)
    sc_chw_to_nhwc_kernel<bf16><<<dim3((HW + 31) / 32, N), 256, 0, (cudaStream_t)stream>>>((const bf16*)dfeat, (bf16*)dact, HW, C,
                                                                                          p_drop, seed, seed_offset);
  else
    sc_chw_to_nhwc_kernel<float><<<dim3((HW + 31) / 32, N), 256, 0, (cudaStream_t)stream>>>((const float*)dfeat, (bf16*)dact, HW, C,
                                                                                           p_drop, seed, seed_offset);
  B2_LAUNCH_CHECK("sc_chw_to_nhwc_kernel");
  return 0;
}
