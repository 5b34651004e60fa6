// Halo-tile tcgen05 convolution for the 3x3 / stride 1 / pad 1 convs of the narrow ResNet stages (C = Cout = 64:
// Bottleneck.conv2 of layer1, BasicBlock convs of layer1; torchvision ResNet under medsos models.py:192).
//
// The im2col-TMA kernel (gemm_tc.cu) fetches every input pixel 9 times (once per filter tap) plus the whole weight
// matrix once per tile: at C = 64 it is bound by L2 -> SM traffic (216 KB per 128-pixel tile, 19 % tensor-pipe
// active in ncu).  Here one tile = TH whole output rows of one image:
//   * ONE TMA (4-D tiled box {64 ch, W+2, TH+2, 1}, out-of-image coordinates zero-filled = the conv padding) brings
//     the (TH+2) x (W+2) halo of input pixels into shared memory, 128 B (64 bf16 channels) per pixel, SWIZZLE_128B;
//   * output pixel (p, q) of the tile gets the row index u = p * (W+2) + q ("padded-width" indexing; the two extra
//     columns per row are computed and discarded, 7 % at W = 28).  Its tap (r, s) input pixel is halo pixel
//     u + r * (W+2) + s: the A operand of tap (r, s) is the SAME shared-memory tile read from a start address
//     shifted by (r * (W+2) + s) * 128 B -- nine descriptors, no data movement;
//   * the 9 x [Cout x 64] weight tiles (72 KB) are loaded once per CTA and stay resident;
//   * the optional BatchNorm+ReLU of the INPUT (BN1 of the block) is applied to the halo tile once per tile by 4
//     transform warps (1.07 passes over the data instead of 9), padding pixels stay zero;
//   * epilogue: raw bf16 output (masked 32-byte row stores), per-channel sum / sum of squares for the following
//     BatchNorm and its in-kernel finalisation by the last CTA.
// L2 -> SM traffic per tile: 23 KB instead of 216 KB.
#include "tc_ptx.cuh"

#include <cudaTypedefs.h>
#include <mutex>

namespace {
using namespace tc;

constexpr int kEpiWarps = 8;
constexpr int kTfWarps = 4;
constexpr int kC = 64;                 // input channels (one 128-byte swizzle row per pixel)
constexpr int kN = 64;                 // output channels
constexpr int kTaps = 9;
constexpr int kStgBytes = 32 * 64;

struct HaloGeom {
  int N, H, W;
  int TH, Wp, Mt;            // output rows per tile, padded width W + 2, TH * Wp (<= 128)
  int tiles_per_img, num_tiles;
  int halo_px;               // (TH + 2) * Wp pixels per halo tile
  int stage_bytes;           // halo_px * 128 rounded up to 1024
  uint32_t magic_wp, magic_tpi;
};

struct InBn {                // relu?(x * scale[c] + shift[c]) on the input halo (null scale: none)
  const float* scale;
  const float* shift;
  int relu;
};

struct OutFin {              // BatchNorm finalisation by the last CTA (null scale: none)
  float* scale;
  float* shift;
  const float* gamma;
  const float* beta;
  float* running_mean;
  float* running_var;
  unsigned int* counter;
  float inv_count, unbias, eps, momentum;
};

__device__ __forceinline__ void tma_load_4d(void* dst, const void* tmap, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// SWIZZLE_128B K-major descriptor whose start address is `rows` 128-byte rows past a 1024-byte-aligned tile base.
// Measured on B200: the tensor core applies the 128-byte swizzle to ABSOLUTE shared-memory address bits (as TMA
// does when it writes the tile), so a row-shifted start needs no matrix-base-offset (setting it to the row phase
// gives wrong operands for every shift that is not a multiple of 8 rows).
__device__ __forceinline__ uint64_t make_sw128_desc_rows(uint32_t tile_base, int rows, int byte_in_row) {
  const uint32_t addr = tile_base + (uint32_t)rows * 128u + (uint32_t)byte_in_row;
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ uint32_t pack_relu(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}

// CS = 64-channel slabs of the input (C = 64 CS), KN = output channels: <.., 1, 64> is the ResNet 64 -> 64 stage, <false, 2, 32>
// the DenseNet dense-layer conv (128 -> growth 32) whose output goes straight into a channel slice of the block buffer (ldy).
template <bool TF, int CS = 1, int KN = 64>
__global__ void __launch_bounds__(64 + 32 * kEpiWarps + (TF ? 32 * kTfWarps : 0), 1)
conv3x3_halo_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                    bf16* __restrict__ y, long ldy, HaloGeom g, InBn at, float* col_sum, float* col_sumsq, OutFin fin) {
  static_assert(!TF || CS == 1, "the input transform is built for one channel slab");
  constexpr int kThreads = 64 + 32 * kEpiWarps + (TF ? 32 * kTfWarps : 0);
  constexpr int kN = KN;
  constexpr int kStages = CS == 1 ? 4 : 3;
  constexpr int kWTileBytes = KN * 128;               // one (tap, slab) [KN x 64] weight tile
  constexpr int kWBytes = kTaps * CS * kWTileBytes;
  constexpr int kActEpi = KN / 32 * 4;                // epilogue warps that own a 32-column chunk
  const int stage_stride = CS * g.stage_bytes;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* w_s = smem;                                        // 9 resident weight tiles
  uint8_t* stage_s = smem + kWBytes;                          // kStages halo tiles (+ 2 KB read slack after the last)
  uint8_t* after = stage_s + kStages * stage_stride + 2048;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(after);
  uint64_t* tf_bar = full_bar + kStages;
  uint64_t* empty_bar = tf_bar + kStages;
  uint64_t* tfull_bar = empty_bar + kStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* w_bar = tempty_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_bar + 1);
  float* stat_s = reinterpret_cast<float*>(after + 256);      // [4 quarters x 2 half-warps][2][64]
  uint8_t* staging_s = after + 256 + 16 * kN * 4;             // kEpiWarps x kStgBytes
  const bool want_stats = col_sum != nullptr;
  for (int i = threadIdx.x; i < 16 * kN; i += kThreads) stat_s[i] = 0.f;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_x);
    prefetch_tmap(&tmap_w);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&tf_bar[i], 32 * kTfWarps);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], kActEpi);
    }
    mbar_init(w_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tc_alloc(tmem_slot, 2 * kN);
    tc_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      mbar_expect_tx(w_bar, kWBytes);
      for (int t = 0; t < kTaps; ++t)
        for (int p = 0; p < CS; ++p)
          tma_load_2d(w_s + (t * CS + p) * kWTileBytes, &tmap_w, w_bar, (t * CS + p) * 64, 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < g.num_tiles; tile += gridDim.x) {
        const int n = g.tiles_per_img == 1 ? tile : (int)__umulhi((uint32_t)tile, g.magic_tpi);
        const int h0 = (tile - n * g.tiles_per_img) * g.TH;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        mbar_expect_tx(&full_bar[stage], (uint32_t)(CS * g.halo_px * 128));
        for (int p = 0; p < CS; ++p)
          tma_load_4d(stage_s + stage * stage_stride + p * g.stage_bytes, &tmap_x, &full_bar[stage], p * 64, -1, h0 - 1, n);
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    constexpr uint32_t idesc = make_idesc(128, kN);
    uint64_t* ready_bar = TF ? tf_bar : full_bar;
    mbar_wait(w_bar, 0);
    // descriptors are built once: the address field (bits 0..13, 16-byte units) of a tap / k-step is the base
    // plus a constant, so the single issuing thread spends two 64-bit adds per MMA instead of two rebuilds
    const uint64_t db0 = make_sw128_desc(smem_u32(w_s));
    uint32_t a_off[kTaps];
#pragma unroll
    for (int t = 0; t < kTaps; ++t) a_off[t] = (uint32_t)(((t / 3) * g.Wp + (t % 3)) * 8);   // halo rows * 128 B / 16
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < g.num_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      mbar_wait(&tempty_bar[acc], ((it >> 1) & 1) ^ 1);
      mbar_wait(&ready_bar[stage], phase);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * kN);
        const uint64_t da0 = make_sw128_desc_rows(smem_u32(stage_s + stage * stage_stride), 0, 0);
#pragma unroll
        for (int t = 0; t < kTaps; ++t) {
#pragma unroll
          for (int p = 0; p < CS; ++p) {
            const uint32_t pa = (uint32_t)(p * (g.stage_bytes / 16));
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              tc_mma_bf16(d_tmem, da0 + (uint64_t)(pa + a_off[t] + k * 2),
                          db0 + (uint64_t)((t * CS + p) * (kWTileBytes / 16) + k * 2), idesc, (t | p | k) != 0);
            }
          }
        }
        tc_commit(&empty_bar[stage]);
        tc_commit(&tfull_bar[acc]);
      }
      __syncwarp();
      if (++stage == kStages) {
        stage = 0;
        phase ^= 1;
      }
    }
  } else if (warp < 2 + kEpiWarps) {
    // =========================== epilogue (warps 2..9) ===========================
    const int quarter = warp & 3;
    const int ch = (warp - 2) >> 2;        // which 32-channel half this warp handles
    uint8_t* stg = staging_s + (warp - 2) * kStgBytes;
    const int sw = (lane >> 1) & 3;
    const int sw_w = lane & 15, sw_hf = lane >> 4;
    int it = 0;
    if (ch < KN / 32)           // (with 32 output channels only the first four epilogue warps have a chunk)
    for (int tile = blockIdx.x; tile < g.num_tiles; tile += gridDim.x, ++it) {
      const int n = g.tiles_per_img == 1 ? tile : (int)__umulhi((uint32_t)tile, g.magic_tpi);
      const int h0 = (tile - n * g.tiles_per_img) * g.TH;
      const int u = quarter * 32 + lane;
      const int pl = (int)__umulhi((uint32_t)u, g.magic_wp);
      const int q = u - pl * g.Wp;
      const bool row_ok = u < g.Mt && q < g.W && h0 + pl < g.H;
      const int acc = it & 1;
      mbar_wait(&tfull_bar[acc], (it >> 1) & 1);
      tc_fence_after();
      uint32_t raw[32];
      tc_ld32(tmem_base + (uint32_t)(acc * kN + ch * 32) + ((uint32_t)(quarter * 32) << 16), raw);
      tc_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);    // accumulator is in registers: release it early
      uint32_t pk[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) pk[j] = pack_bf16x2(__uint_as_float(raw[2 * j]), __uint_as_float(raw[2 * j + 1]));
      if (row_ok) {
        bf16* dp = y + ((((long)n * g.H + h0 + pl) * g.W + q) * ldy + ch * 32);
        asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dp), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]),
                     "r"(pk[3]), "r"(pk[4]), "r"(pk[5]), "r"(pk[6]), "r"(pk[7])
                     : "memory");
        asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dp + 16), "r"(pk[8]), "r"(pk[9]),
                     "r"(pk[10]), "r"(pk[11]), "r"(pk[12]), "r"(pk[13]), "r"(pk[14]), "r"(pk[15])
                     : "memory");
      }
      if (want_stats) {
        __syncwarp();
#pragma unroll
        for (int c = 0; c < 4; ++c)
          *reinterpret_cast<uint4*>(stg + lane * 64 + ((c ^ sw) << 4)) =
              row_ok ? make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]) : make_uint4(0u, 0u, 0u, 0u);
        __syncwarp();
        float2 s1a = make_float2(0.f, 0.f), s1b = s1a, s2a = s1a, s2b = s1a;
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
          const int k = (i >> 1) & 3;
          const uint32_t col = (uint32_t)((((sw_w >> 2) ^ k) << 4) + (sw_w & 3) * 4) + (uint32_t)(sw_hf * 1024);
          const uint32_t u0 = *reinterpret_cast<const uint32_t*>(stg + col + (uint32_t)(sw_hf * 64) + i * 64);
          const uint32_t u1 = *reinterpret_cast<const uint32_t*>(stg + col - (uint32_t)(sw_hf * 64) + (i + 1) * 64);
          const float2 x0 = make_float2(__uint_as_float(u0 << 16), __uint_as_float(u0 & 0xffff0000u));
          const float2 x1 = make_float2(__uint_as_float(u1 << 16), __uint_as_float(u1 & 0xffff0000u));
          s1a = __fadd2_rn(s1a, x0);
          s1b = __fadd2_rn(s1b, x1);
          s2a = __ffma2_rn(x0, x0, s2a);
          s2b = __ffma2_rn(x1, x1, s2b);
        }
        float2* st = reinterpret_cast<float2*>(stat_s + (quarter * 2 + sw_hf) * 2 * kN + ch * 32 + 2 * sw_w);
        float2* st2 = reinterpret_cast<float2*>(stat_s + (quarter * 2 + sw_hf) * 2 * kN + kN + ch * 32 + 2 * sw_w);
        *st = __fadd2_rn(*st, __fadd2_rn(s1a, s1b));
        *st2 = __fadd2_rn(*st2, __fadd2_rn(s2a, s2b));
      }
    }
  } else if (TF) {
    // =========================== input BatchNorm + ReLU on the halo tile (warps 10..13) ===========================
    const int tt = threadIdx.x - (64 + 32 * kEpiWarps);
    const int c = tt & 7;              // 16-byte chunk of the pixel's 128-byte row: channels 8c .. 8c+7
    const int rb = tt >> 3;            // pixels rb, rb + 16, ...
    const float4 sc0 = __ldg(reinterpret_cast<const float4*>(at.scale + c * 8));
    const float4 sc1 = __ldg(reinterpret_cast<const float4*>(at.scale + c * 8 + 4));
    const float4 sh0 = __ldg(reinterpret_cast<const float4*>(at.shift + c * 8));
    const float4 sh1 = __ldg(reinterpret_cast<const float4*>(at.shift + c * 8 + 4));
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < g.num_tiles; tile += gridDim.x) {
      const int n = g.tiles_per_img == 1 ? tile : (int)__umulhi((uint32_t)tile, g.magic_tpi);
      const int h0 = (tile - n * g.tiles_per_img) * g.TH;
      mbar_wait(&full_bar[stage], phase);
      uint8_t* sa = stage_s + stage * stage_stride;
      for (int i0 = rb; i0 < g.halo_px; i0 += 16 * 4) {
        uint4 u[4];
        bool ok[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int i = i0 + 16 * j;
          const int hy = (int)__umulhi((uint32_t)i, g.magic_wp);
          const int wx = i - hy * g.Wp;
          ok[j] = i < g.halo_px && (unsigned)(h0 - 1 + hy) < (unsigned)g.H && wx >= 1 && wx <= g.W;
          if (i < g.halo_px) u[j] = *reinterpret_cast<const uint4*>(sa + i * 128 + ((c ^ (i & 7)) << 4));
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int i = i0 + 16 * j;
          if (i >= g.halo_px) continue;
          float2 a0 = make_float2(__uint_as_float(u[j].x << 16), __uint_as_float(u[j].x & 0xffff0000u));
          float2 a1 = make_float2(__uint_as_float(u[j].y << 16), __uint_as_float(u[j].y & 0xffff0000u));
          float2 a2 = make_float2(__uint_as_float(u[j].z << 16), __uint_as_float(u[j].z & 0xffff0000u));
          float2 a3 = make_float2(__uint_as_float(u[j].w << 16), __uint_as_float(u[j].w & 0xffff0000u));
          a0 = __ffma2_rn(a0, make_float2(sc0.x, sc0.y), make_float2(sh0.x, sh0.y));
          a1 = __ffma2_rn(a1, make_float2(sc0.z, sc0.w), make_float2(sh0.z, sh0.w));
          a2 = __ffma2_rn(a2, make_float2(sc1.x, sc1.y), make_float2(sh1.x, sh1.y));
          a3 = __ffma2_rn(a3, make_float2(sc1.z, sc1.w), make_float2(sh1.z, sh1.w));
          uint4 t;
          if (at.relu) {
            t.x = pack_relu(a0.x, a0.y); t.y = pack_relu(a1.x, a1.y); t.z = pack_relu(a2.x, a2.y); t.w = pack_relu(a3.x, a3.y);
          } else {
            t.x = pack_bf16x2(a0.x, a0.y); t.y = pack_bf16x2(a1.x, a1.y); t.z = pack_bf16x2(a2.x, a2.y); t.w = pack_bf16x2(a3.x, a3.y);
          }
          if (!ok[j]) t = make_uint4(0u, 0u, 0u, 0u);        // conv padding / out-of-image rows stay zero
          *reinterpret_cast<uint4*>(sa + i * 128 + ((c ^ (i & 7)) << 4)) = t;
        }
      }
      fence_proxy_async_smem();
      mbar_arrive(&tf_bar[stage]);
      if (++stage == kStages) {
        stage = 0;
        phase ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (want_stats) {
    for (int c = threadIdx.x; c < kN; c += kThreads) {
      float a1 = 0.f, a2 = 0.f;
#pragma unroll
      for (int qq = 0; qq < 8; ++qq) {
        a1 += stat_s[qq * 2 * kN + c];
        a2 += stat_s[qq * 2 * kN + kN + c];
      }
      atomicAdd(col_sum + c, a1);
      atomicAdd(col_sumsq + c, a2);
    }
  }
  if (fin.scale != nullptr) {
    __shared__ int is_last;
    __threadfence();          // this thread's statistics reductions are device-visible before the block barrier
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(fin.counter, 1u) == gridDim.x - 1);
    __syncthreads();
    if (is_last) {
      __threadfence();
      for (int c = threadIdx.x; c < kN; c += kThreads) {
        const float mean = __ldcg(col_sum + c) * fin.inv_count;
        const float var = fmaxf(__ldcg(col_sumsq + c) * fin.inv_count - mean * mean, 0.f);
        if (fin.running_mean != nullptr) {
          fin.running_mean[c] = (1.f - fin.momentum) * fin.running_mean[c] + fin.momentum * mean;
          fin.running_var[c] = (1.f - fin.momentum) * fin.running_var[c] + fin.momentum * var * fin.unbias;
        }
        const float sc = fin.gamma[c] * rsqrtf(var + fin.eps);
        fin.scale[c] = sc;
        fin.shift[c] = fin.beta[c] - mean * sc;
      }
    }
  }
  if (warp == 1) {
    tc_fence_after();
    tc_dealloc(tmem_base, 2 * kN);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// 128 -> 128 channel stage (Bottleneck.conv2 of layer2; BasicBlock convs of layer2): the 288 KB of weights cannot stay
// resident, so they STREAM through a ring of [128 x 64] (tap, slab) tiles -- the same 18 tiles in the same order for every
// CTA, which L2 serves once per wave (identical requests from neighbouring SMs are de-duplicated) -- and every weight
// tile is used by TWO 128-pixel MMA tiles of one halo: a stage holds the halo of 2 TH output rows (the whole 14x14 image at
// layer2 of a 112x112 frame: 16 x 16 pixels x 2 slabs = 64 KB), the accumulators are 2 x 128 columns per set, two sets.
// L2 -> SM per 256 output pixels: 64 KB of halo + 288 KB of (shared) weights instead of 2 x 590 KB through the im2col form.
// BN1 + ReLU of the input is applied to the halo in shared memory (no stand-alone pass over the tensor).
constexpr int kSN = 128;               // output channels
constexpr int kSCS = 2;                // input slabs (128 channels)
constexpr int kSWTile = kSN * 128;     // 16 KB: [128 x 64] weight tile
constexpr int kSWTiles = kTaps * kSCS; // 18 per super tile

struct StreamGeom {
  int N, H, W;
  int TH, Wp;                // rows per 128-pixel MMA tile, padded width
  int rows_super;            // 2 TH output rows per stage
  int box_rows;              // halo rows fetched per stage (<= rows_super + 2; rows only discarded outputs need are left out)
  int tiles_per_img, num_tiles;
  int halo_px, slab_bytes;   // box_rows * Wp pixels; 128 B each, rounded up to 1024
  int w_stages;
  uint32_t magic_wp, magic_tpi;
};

template <bool TF>
__global__ void __launch_bounds__(96 + 32 * kEpiWarps + (TF ? 32 * kTfWarps : 0), 1)
conv3x3_halo_stream_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                           bf16* __restrict__ y, StreamGeom g, InBn at, float* col_sum, float* col_sumsq, OutFin fin) {
  constexpr int kThreads = 96 + 32 * kEpiWarps + (TF ? 32 * kTfWarps : 0);
  constexpr int kStages = 2;
  constexpr int kMaxW = 4;
  const int stage_stride = kSCS * g.slab_bytes;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* stage_s = smem;                                    // kStages halo tiles
  uint8_t* w_s = smem + kStages * stage_stride;               // weight ring (also the read slack of the last halo slab)
  uint8_t* after = w_s + g.w_stages * kSWTile;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(after);
  uint64_t* tf_bar = full_bar + kStages;
  uint64_t* empty_bar = tf_bar + kStages;
  uint64_t* wfull_bar = empty_bar + kStages;
  uint64_t* wempty_bar = wfull_bar + kMaxW;
  uint64_t* tfull_bar = wempty_bar + kMaxW;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  float* stat_s = reinterpret_cast<float*>(after + 256);      // [4 quarters x 2 half-warps][2][128]
  uint8_t* staging_s = after + 256 + 16 * kSN * 4;            // kEpiWarps x kStgBytes
  const bool want_stats = col_sum != nullptr;
  for (int i = threadIdx.x; i < 16 * kSN; i += kThreads) stat_s[i] = 0.f;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_x);
    prefetch_tmap(&tmap_w);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&tf_bar[i], 32 * kTfWarps);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < kMaxW; ++i) {
      mbar_init(&wfull_bar[i], 1);
      mbar_init(&wempty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], kEpiWarps);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tc_alloc(tmem_slot, 512);
    tc_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // =========================== halo producer ===========================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < g.num_tiles; tile += gridDim.x) {
        const int n = g.tiles_per_img == 1 ? tile : (int)__umulhi((uint32_t)tile, g.magic_tpi);
        const int h0 = (tile - n * g.tiles_per_img) * g.rows_super;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        mbar_expect_tx(&full_bar[stage], (uint32_t)(kSCS * g.halo_px * 128));
        for (int p = 0; p < kSCS; ++p)
          tma_load_4d(stage_s + stage * stage_stride + p * g.slab_bytes, &tmap_x, &full_bar[stage], p * 64, -1, h0 - 1, n);
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 2) {
    // =========================== weight producer: 18 tiles per super tile, the same order for everybody ===========================
    if (lane == 0) {
      int ws = 0;
      uint32_t wphase = 0;
      for (int tile = blockIdx.x; tile < g.num_tiles; tile += gridDim.x) {
        for (int j = 0; j < kSWTiles; ++j) {
          mbar_wait(&wempty_bar[ws], wphase ^ 1);
          mbar_expect_tx(&wfull_bar[ws], kSWTile);
          tma_load_2d(w_s + ws * kSWTile, &tmap_w, &wfull_bar[ws], j * 64, 0);
          if (++ws == g.w_stages) {
            ws = 0;
            wphase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    constexpr uint32_t idesc = make_idesc(128, kSN);
    uint64_t* ready_bar = TF ? tf_bar : full_bar;
    const uint64_t db0 = make_sw128_desc(smem_u32(w_s));
    const uint32_t m_off = (uint32_t)(g.TH * g.Wp * 8);        // second MMA tile: TH rows further (halo rows * 128 B / 16)
    int stage = 0, ws = 0, it = 0;
    uint32_t phase = 0, wphase = 0;
    for (int tile = blockIdx.x; tile < g.num_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      mbar_wait(&tempty_bar[acc], ((it >> 1) & 1) ^ 1);
      mbar_wait(&ready_bar[stage], phase);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 2 * kSN);
      const uint64_t da0 = make_sw128_desc_rows(smem_u32(stage_s + stage * stage_stride), 0, 0);
#pragma unroll 1
      for (int t = 0; t < kTaps; ++t) {
        const uint32_t a_off = (uint32_t)(((t / 3) * g.Wp + (t % 3)) * 8);
#pragma unroll
        for (int p = 0; p < kSCS; ++p) {
          mbar_wait(&wfull_bar[ws], wphase);
          tc_fence_after();
          if (lane == 0) {
            const uint32_t pa = (uint32_t)(p * (g.slab_bytes / 16)) + a_off;
            const uint64_t db = db0 + (uint64_t)(ws * (kSWTile / 16));
#pragma unroll
            for (int m = 0; m < 2; ++m) {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                tc_mma_bf16(d_tmem + (uint32_t)(m * kSN), da0 + (uint64_t)(pa + m * m_off + k * 2), db + (uint64_t)(k * 2), idesc,
                            (t | p | k) != 0);
            }
            tc_commit(&wempty_bar[ws]);
          }
          __syncwarp();
          if (++ws == g.w_stages) {
            ws = 0;
            wphase ^= 1;
          }
        }
      }
      if (lane == 0) {
        tc_commit(&empty_bar[stage]);
        tc_commit(&tfull_bar[acc]);
      }
      __syncwarp();
      if (++stage == kStages) {
        stage = 0;
        phase ^= 1;
      }
    }
  } else if (warp < 3 + kEpiWarps) {
    // =========================== epilogue (warps 3..10): quarter = warp % 4, two 32-column chunks x two MMA tiles each ==========
    const int quarter = warp & 3;
    const int cg = (warp - 3) >> 2;        // 64-column group
    uint8_t* stg = staging_s + (warp - 3) * kStgBytes;
    const int sw = (lane >> 1) & 3;
    const int sw_w = lane & 15, sw_hf = lane >> 4;
    int it = 0;
    for (int tile = blockIdx.x; tile < g.num_tiles; tile += gridDim.x, ++it) {
      const int n = g.tiles_per_img == 1 ? tile : (int)__umulhi((uint32_t)tile, g.magic_tpi);
      const int h0 = (tile - n * g.tiles_per_img) * g.rows_super;
      const int acc = it & 1;
      mbar_wait(&tfull_bar[acc], (it >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int m = 0; m < 2; ++m) {
        const int u = quarter * 32 + lane;                     // row of the MMA tile
        const int pl = (int)__umulhi((uint32_t)u, g.magic_wp);
        const int q = u - pl * g.Wp;
        const int prow = h0 + m * g.TH + pl;
        const bool row_ok = pl < g.TH && q < g.W && prow < g.H;
#pragma unroll 1
        for (int cc = 0; cc < 2; ++cc) {
          const int ch = cg * 2 + cc;                          // 32-column chunk
          uint32_t raw[32];
          tc_ld32(tmem_base + (uint32_t)(acc * 2 * kSN + m * kSN + ch * 32) + ((uint32_t)(quarter * 32) << 16), raw);
          tc_wait_ld();
          if (m == 1 && cc == 1) {                             // the whole accumulator set is in registers / written
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[acc]);
          }
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) pk[j] = pack_bf16x2(__uint_as_float(raw[2 * j]), __uint_as_float(raw[2 * j + 1]));
          if (row_ok) {
            bf16* dp = y + ((((long)n * g.H + prow) * g.W + q) * kSN + ch * 32);
            asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dp), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]),
                         "r"(pk[3]), "r"(pk[4]), "r"(pk[5]), "r"(pk[6]), "r"(pk[7])
                         : "memory");
            asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dp + 16), "r"(pk[8]), "r"(pk[9]),
                         "r"(pk[10]), "r"(pk[11]), "r"(pk[12]), "r"(pk[13]), "r"(pk[14]), "r"(pk[15])
                         : "memory");
          }
          if (want_stats) {
            __syncwarp();
#pragma unroll
            for (int c = 0; c < 4; ++c)
              *reinterpret_cast<uint4*>(stg + lane * 64 + ((c ^ sw) << 4)) =
                  row_ok ? make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]) : make_uint4(0u, 0u, 0u, 0u);
            __syncwarp();
            float2 s1a = make_float2(0.f, 0.f), s1b = s1a, s2a = s1a, s2b = s1a;
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
              const int k = (i >> 1) & 3;
              const uint32_t col = (uint32_t)((((sw_w >> 2) ^ k) << 4) + (sw_w & 3) * 4) + (uint32_t)(sw_hf * 1024);
              const uint32_t u0 = *reinterpret_cast<const uint32_t*>(stg + col + (uint32_t)(sw_hf * 64) + i * 64);
              const uint32_t u1 = *reinterpret_cast<const uint32_t*>(stg + col - (uint32_t)(sw_hf * 64) + (i + 1) * 64);
              const float2 x0 = make_float2(__uint_as_float(u0 << 16), __uint_as_float(u0 & 0xffff0000u));
              const float2 x1 = make_float2(__uint_as_float(u1 << 16), __uint_as_float(u1 & 0xffff0000u));
              s1a = __fadd2_rn(s1a, x0);
              s1b = __fadd2_rn(s1b, x1);
              s2a = __ffma2_rn(x0, x0, s2a);
              s2b = __ffma2_rn(x1, x1, s2b);
            }
            float2* st = reinterpret_cast<float2*>(stat_s + (quarter * 2 + sw_hf) * 2 * kSN + ch * 32 + 2 * sw_w);
            float2* st2 = reinterpret_cast<float2*>(stat_s + (quarter * 2 + sw_hf) * 2 * kSN + kSN + ch * 32 + 2 * sw_w);
            *st = __fadd2_rn(*st, __fadd2_rn(s1a, s1b));
            *st2 = __fadd2_rn(*st2, __fadd2_rn(s2a, s2b));
            __syncwarp();
          }
        }
      }
    }
  } else if (TF) {
    // =========================== input BatchNorm + ReLU on the halo (warps 11..14) ===========================
    const int tt = threadIdx.x - (96 + 32 * kEpiWarps);
    const int c = tt & 7;              // 16-byte chunk of the pixel's 128-byte row: channels 8c .. 8c+7 of a slab
    const int rb = tt >> 3;            // pixels rb, rb + 16, ...
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < g.num_tiles; tile += gridDim.x) {
      const int n = g.tiles_per_img == 1 ? tile : (int)__umulhi((uint32_t)tile, g.magic_tpi);
      const int h0 = (tile - n * g.tiles_per_img) * g.rows_super;
      mbar_wait(&full_bar[stage], phase);
#pragma unroll 1
      for (int p = 0; p < kSCS; ++p) {
        const float4 sc0 = __ldg(reinterpret_cast<const float4*>(at.scale + p * 64 + c * 8));
        const float4 sc1 = __ldg(reinterpret_cast<const float4*>(at.scale + p * 64 + c * 8 + 4));
        const float4 sh0 = __ldg(reinterpret_cast<const float4*>(at.shift + p * 64 + c * 8));
        const float4 sh1 = __ldg(reinterpret_cast<const float4*>(at.shift + p * 64 + c * 8 + 4));
        uint8_t* sa = stage_s + stage * stage_stride + p * g.slab_bytes;
        for (int i0 = rb; i0 < g.halo_px; i0 += 16 * 4) {
          uint4 uu[4];
          bool ok[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int i = i0 + 16 * j;
            const int hy = (int)__umulhi((uint32_t)i, g.magic_wp);
            const int wx = i - hy * g.Wp;
            ok[j] = i < g.halo_px && (unsigned)(h0 - 1 + hy) < (unsigned)g.H && wx >= 1 && wx <= g.W;
            if (i < g.halo_px) uu[j] = *reinterpret_cast<const uint4*>(sa + i * 128 + ((c ^ (i & 7)) << 4));
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int i = i0 + 16 * j;
            if (i >= g.halo_px) continue;
            float2 a0 = make_float2(__uint_as_float(uu[j].x << 16), __uint_as_float(uu[j].x & 0xffff0000u));
            float2 a1 = make_float2(__uint_as_float(uu[j].y << 16), __uint_as_float(uu[j].y & 0xffff0000u));
            float2 a2 = make_float2(__uint_as_float(uu[j].z << 16), __uint_as_float(uu[j].z & 0xffff0000u));
            float2 a3 = make_float2(__uint_as_float(uu[j].w << 16), __uint_as_float(uu[j].w & 0xffff0000u));
            a0 = __ffma2_rn(a0, make_float2(sc0.x, sc0.y), make_float2(sh0.x, sh0.y));
            a1 = __ffma2_rn(a1, make_float2(sc0.z, sc0.w), make_float2(sh0.z, sh0.w));
            a2 = __ffma2_rn(a2, make_float2(sc1.x, sc1.y), make_float2(sh1.x, sh1.y));
            a3 = __ffma2_rn(a3, make_float2(sc1.z, sc1.w), make_float2(sh1.z, sh1.w));
            uint4 t;
            if (at.relu) {
              t.x = pack_relu(a0.x, a0.y); t.y = pack_relu(a1.x, a1.y); t.z = pack_relu(a2.x, a2.y); t.w = pack_relu(a3.x, a3.y);
            } else {
              t.x = pack_bf16x2(a0.x, a0.y); t.y = pack_bf16x2(a1.x, a1.y); t.z = pack_bf16x2(a2.x, a2.y); t.w = pack_bf16x2(a3.x, a3.y);
            }
            if (!ok[j]) t = make_uint4(0u, 0u, 0u, 0u);        // conv padding / out-of-image rows stay zero
            *reinterpret_cast<uint4*>(sa + i * 128 + ((c ^ (i & 7)) << 4)) = t;
          }
        }
      }
      fence_proxy_async_smem();
      mbar_arrive(&tf_bar[stage]);
      if (++stage == kStages) {
        stage = 0;
        phase ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (want_stats) {
    for (int c = threadIdx.x; c < kSN; c += kThreads) {
      float a1 = 0.f, a2 = 0.f;
#pragma unroll
      for (int qq = 0; qq < 8; ++qq) {
        a1 += stat_s[qq * 2 * kSN + c];
        a2 += stat_s[qq * 2 * kSN + kSN + c];
      }
      atomicAdd(col_sum + c, a1);
      atomicAdd(col_sumsq + c, a2);
    }
  }
  if (fin.scale != nullptr) {
    __shared__ int is_last;
    __threadfence();          // this thread's statistics reductions are device-visible before the block barrier
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(fin.counter, 1u) == gridDim.x - 1);
    __syncthreads();
    if (is_last) {
      __threadfence();
      for (int c = threadIdx.x; c < kSN; c += kThreads) {
        const float mean = __ldcg(col_sum + c) * fin.inv_count;
        const float var = fmaxf(__ldcg(col_sumsq + c) * fin.inv_count - mean * mean, 0.f);
        if (fin.running_mean != nullptr) {
          fin.running_mean[c] = (1.f - fin.momentum) * fin.running_mean[c] + fin.momentum * mean;
          fin.running_var[c] = (1.f - fin.momentum) * fin.running_var[c] + fin.momentum * var * fin.unbias;
        }
        const float sc = fin.gamma[c] * rsqrtf(var + fin.eps);
        fin.scale[c] = sc;
        fin.shift[c] = fin.beta[c] - mean * sc;
      }
    }
  }
  if (warp == 1) {
    tc_fence_after();
    tc_dealloc(tmem_base, 512);
  }
}

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

int make_geom(HaloGeom* g, int N, int H, int W) {
  g->N = N; g->H = H; g->W = W;
  g->Wp = W + 2;
  if (g->Wp > 64) return -1;
  g->TH = 128 / g->Wp;
  if (g->TH > H) g->TH = H;
  g->Mt = g->TH * g->Wp;
  g->tiles_per_img = (H + g->TH - 1) / g->TH;
  const long tiles = (long)N * g->tiles_per_img;
  if (tiles >= (1L << 31) || (unsigned long)tiles * g->tiles_per_img >= (1ul << 32)) return -1;
  g->num_tiles = (int)tiles;
  g->halo_px = (g->TH + 2) * g->Wp;
  g->stage_bytes = (g->halo_px * 128 + 1023) / 1024 * 1024;
  g->magic_wp = (uint32_t)(((1ull << 32) + g->Wp - 1) / g->Wp);
  g->magic_tpi = g->tiles_per_img == 1 ? 0u : (uint32_t)(((1ull << 32) + g->tiles_per_img - 1) / g->tiles_per_img);
  return 0;
}

int halo_smem_bytes(const HaloGeom& g, int cs = 1, int kn = kN) {
  const int stages = cs == 1 ? 4 : 3;
  // the shifted A descriptors of the last taps read up to 128 + 2 Wp + 2 halo rows from a panel start: with small images
  // (few halo rows per tile) that runs past the panel, the following stages and the tail regions -- harmless garbage (it only
  // feeds discarded accumulator rows) but it must stay inside the CTA's allocation, so the tail is padded accordingly
  const int reach = (128 + 2 * g.Wp + 3) * 128;
  const int tail = 2048 + 256 + 16 * kn * 4 + (kn / 32 * 4) * kStgBytes + 1024;
  const int pad = reach > g.stage_bytes + tail ? reach - g.stage_bytes - tail : 0;
  return kTaps * cs * kn * 128 + stages * cs * g.stage_bytes + tail + pad;
}

int make_stream_geom(StreamGeom* g, int N, int H, int W) {
  g->N = N; g->H = H; g->W = W;
  g->Wp = W + 2;
  if (g->Wp > 64) return -1;
  g->TH = 128 / g->Wp;
  g->rows_super = 2 * g->TH;
  g->box_rows = (g->rows_super < H ? g->rows_super : H) + 2;
  g->tiles_per_img = (H + g->rows_super - 1) / g->rows_super;
  const long tiles = (long)N * g->tiles_per_img;
  if (tiles >= (1L << 31) || (unsigned long)tiles * g->tiles_per_img >= (1ul << 32)) return -1;
  g->num_tiles = (int)tiles;
  g->halo_px = g->box_rows * g->Wp;
  g->slab_bytes = (g->halo_px * 128 + 1023) / 1024 * 1024;
  g->magic_wp = (uint32_t)(((1ull << 32) + g->Wp - 1) / g->Wp);
  g->magic_tpi = g->tiles_per_img == 1 ? 0u : (uint32_t)(((1ull << 32) + g->tiles_per_img - 1) / g->tiles_per_img);
  g->w_stages = 0;
  for (int ws = 4; ws >= 3; --ws) {
    const int total = 2 * kSCS * g->slab_bytes + ws * kSWTile + 256 + 16 * kSN * 4 + kEpiWarps * kStgBytes + 1024;
    // the shifted descriptors of the second MMA tile read up to (TH Wp + 128 + 2 Wp + 3) pixel rows from a slab start
    // (garbage that only feeds discarded accumulator rows): from the LAST slab that must stay inside the allocation
    const int reach = (g->TH * g->Wp + 128 + 2 * g->Wp + 3) * 128;
    if (total <= 227 * 1024 && reach <= g->slab_bytes + ws * kSWTile + 256 + 16 * kSN * 4 + kEpiWarps * kStgBytes) {
      g->w_stages = ws;
      break;
    }
  }
  return g->w_stages ? 0 : -1;
}

int stream_smem_bytes(const StreamGeom& g) {
  return 2 * kSCS * g.slab_bytes + g.w_stages * kSWTile + 256 + 16 * kSN * 4 + kEpiWarps * kStgBytes + 1024;
}

}  // namespace

// 1 when b2_conv3x3_halo_bn_nhwc_bf16 supports the shape (3x3 / stride 1 / pad 1; C = Cout = 64 with resident weights or
// C = Cout = 128 with streamed weights; W <= 62 and the halo stages must fit shared memory)
B2_API int b2_conv3x3_halo_supported(int N, int H, int W, int C, int Cout) {
  if (N <= 0 || H <= 0 || W <= 0) return 0;
  if (C == 128 && Cout == 128) {
    static const bool off = getenv("B2_NO_HALO128") != nullptr;
    StreamGeom sg;
    return !off && make_stream_geom(&sg, N, H, W) == 0 ? 1 : 0;
  }
  HaloGeom g;
  if (C != kC || Cout != kN) return 0;
  if (make_geom(&g, N, H, W) != 0) return 0;
  return halo_smem_bytes(g) <= 227 * 1024 ? 1 : 0;
}

// y [N,H,W,64] = conv3x3(relu?(x * a_scale + a_shift)) (stride 1, pad 1), raw bf16 + statistics + finalisation ;
// see include/b200lrcn.h
B2_API int b2_conv3x3_halo_bn_nhwc_bf16(const void* x, int N, int H, int W, int C, const void* w, int Cout, void* y,
                                        const float* a_scale, const float* a_shift, int a_relu, float* col_sum,
                                        float* col_sumsq, const float* fin_gamma, const float* fin_beta,
                                        float* fin_running_mean, float* fin_running_var, float* fin_scale,
                                        float* fin_shift, unsigned int* fin_counter, float eps, float momentum,
                                        void* stream) {
  const char* who = "b2_conv3x3_halo_bn_nhwc_bf16";
  B2_ARG_CHECK(x && w && y, "%s: null pointer", who);
  B2_ARG_CHECK(b2_conv3x3_halo_supported(N, H, W, C, Cout), "%s: unsupported shape N=%d H=%d W=%d C=%d Cout=%d", who, N, H,
               W, C, Cout);
  B2_ARG_CHECK((a_scale == nullptr) == (a_shift == nullptr), "%s: a_scale and a_shift go together", who);
  B2_ARG_CHECK((col_sum == nullptr) == (col_sumsq == nullptr), "%s: col_sum and col_sumsq go together", who);
  B2_ARG_CHECK(fin_scale == nullptr || (fin_shift && fin_gamma && fin_beta && fin_counter && col_sum),
               "%s: BatchNorm finalisation needs gamma/beta/shift/counter and the statistics buffers", who);
  B2_ARG_CHECK(((uintptr_t)x & 15) == 0 && ((uintptr_t)w & 15) == 0 && ((uintptr_t)y & 31) == 0,
               "%s: x / w must be 16 B and y 32 B aligned", who);
  if (int r = load_encode()) return r;
  if (C == 128) {
    StreamGeom sg;
    make_stream_geom(&sg, N, H, W);
    CUtensorMap tx, tw;
    {
      cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
      cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
      cuuint32_t box[4] = {64u, (cuuint32_t)sg.Wp, (cuuint32_t)sg.box_rows, 1};
      cuuint32_t estr[4] = {1, 1, 1, 1};
      CUresult cr = g_encode(&tx, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), dims, strides, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (cr != CUDA_SUCCESS) {
        b2_set_error("%s: cuTensorMapEncodeTiled(x) failed (%d)", who, (int)cr);
        return -3;
      }
    }
    {
      cuuint64_t dims[2] = {(cuuint64_t)(kTaps * C), (cuuint64_t)Cout};
      cuuint64_t strides[1] = {(cuuint64_t)(kTaps * C) * 2};
      cuuint32_t box[2] = {64u, (cuuint32_t)kSN};
      cuuint32_t estr[2] = {1, 1};
      CUresult cr = g_encode(&tw, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w), dims, strides, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (cr != CUDA_SUCCESS) {
        b2_set_error("%s: cuTensorMapEncodeTiled(w) failed (%d)", who, (int)cr);
        return -3;
      }
    }
    InBn at = {a_scale, a_shift, a_relu};
    OutFin fin = {};
    if (fin_scale != nullptr) {
      const double count = (double)N * H * W;
      fin.scale = fin_scale;
      fin.shift = fin_shift;
      fin.gamma = fin_gamma;
      fin.beta = fin_beta;
      fin.running_mean = fin_running_mean;
      fin.running_var = fin_running_var;
      fin.counter = fin_counter;
      fin.inv_count = (float)(1.0 / count);
      fin.unbias = count > 1 ? (float)(count / (count - 1.0)) : 1.f;
      fin.eps = eps;
      fin.momentum = momentum;
    }
    const int smem = stream_smem_bytes(sg);
    static B2PerDeviceMax attr_smem[2];
    const int tf = a_scale != nullptr;
    if (attr_smem[tf].below(smem)) {
      if (tf)
        B2_CUDA_CHECK(cudaFuncSetAttribute(conv3x3_halo_stream_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      else
        B2_CUDA_CHECK(cudaFuncSetAttribute(conv3x3_halo_stream_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      attr_smem[tf].set(smem);
    }
    const int grid = sg.num_tiles < b2_num_sms() ? sg.num_tiles : b2_num_sms();
    cudaStream_t st = (cudaStream_t)stream;
    if (tf)
      conv3x3_halo_stream_kernel<true><<<grid, 96 + 32 * kEpiWarps + 32 * kTfWarps, smem, st>>>(tx, tw, (bf16*)y, sg, at, col_sum,
                                                                                               col_sumsq, fin);
    else
      conv3x3_halo_stream_kernel<false><<<grid, 96 + 32 * kEpiWarps, smem, st>>>(tx, tw, (bf16*)y, sg, at, col_sum, col_sumsq, fin);
    B2_LAUNCH_CHECK("conv3x3_halo_stream_kernel");
    return 0;
  }
  HaloGeom g;
  make_geom(&g, N, H, W);
  CUtensorMap tx, tw;
  {
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {(cuuint32_t)kC, (cuuint32_t)g.Wp, (cuuint32_t)(g.TH + 2), 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult cr = g_encode(&tx, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) {
      b2_set_error("%s: cuTensorMapEncodeTiled(x) failed (%d)", who, (int)cr);
      return -3;
    }
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)(kTaps * C), (cuuint64_t)Cout};
    cuuint64_t strides[1] = {(cuuint64_t)(kTaps * C) * 2};
    cuuint32_t box[2] = {(cuuint32_t)kC, (cuuint32_t)kN};
    cuuint32_t estr[2] = {1, 1};
    CUresult cr = g_encode(&tw, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) {
      b2_set_error("%s: cuTensorMapEncodeTiled(w) failed (%d)", who, (int)cr);
      return -3;
    }
  }
  InBn at = {a_scale, a_shift, a_relu};
  OutFin fin = {};
  if (fin_scale != nullptr) {
    const double count = (double)N * H * W;
    fin.scale = fin_scale;
    fin.shift = fin_shift;
    fin.gamma = fin_gamma;
    fin.beta = fin_beta;
    fin.running_mean = fin_running_mean;
    fin.running_var = fin_running_var;
    fin.counter = fin_counter;
    fin.inv_count = (float)(1.0 / count);
    fin.unbias = count > 1 ? (float)(count / (count - 1.0)) : 1.f;
    fin.eps = eps;
    fin.momentum = momentum;
  }
  const int smem = halo_smem_bytes(g);
  static B2PerDeviceMax attr_smem[2];
  const int tf = a_scale != nullptr;
  if (attr_smem[tf].below(smem)) {
    if (tf)
      B2_CUDA_CHECK(cudaFuncSetAttribute(conv3x3_halo_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    else
      B2_CUDA_CHECK(cudaFuncSetAttribute(conv3x3_halo_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_smem[tf].set(smem);
  }
  const int grid = g.num_tiles < b2_num_sms() ? g.num_tiles : b2_num_sms();
  cudaStream_t st = (cudaStream_t)stream;
  if (tf)
    conv3x3_halo_kernel<true><<<grid, 64 + 32 * kEpiWarps + 32 * kTfWarps, smem, st>>>(tx, tw, (bf16*)y, (long)kN, g, at,
                                                                                      col_sum, col_sumsq, fin);
  else
    conv3x3_halo_kernel<false><<<grid, 64 + 32 * kEpiWarps, smem, st>>>(tx, tw, (bf16*)y, (long)kN, g, at, col_sum, col_sumsq,
                                                                       fin);
  B2_LAUNCH_CHECK("conv3x3_halo_kernel");
  return 0;
}

// ---- DenseNet dense-layer conv: y[:, :32] (row stride ldy: a channel slice of the block buffer) = conv3x3(x [N,H,W,128]),
// stride 1, pad 1, + per-channel statistics of the output (csrc/dense_ops.cu explains the buffer layout)
B2_API int b2_conv3x3_halo_dense_supported(int N, int H, int W, int C, int Cout) {
  HaloGeom g;
  if (C != 128 || Cout != 32 || N <= 0 || H <= 0 || W <= 0) return 0;
  if (make_geom(&g, N, H, W) != 0) return 0;
  return halo_smem_bytes(g, 2, 32) <= 227 * 1024 ? 1 : 0;
}

B2_API int b2_conv3x3_halo_dense_bf16(const void* x, int N, int H, int W, int C, const void* w, int Cout, void* y, long ldy,
                                      float* col_sum, float* col_sumsq, void* stream) {
  const char* who = "b2_conv3x3_halo_dense_bf16";
  B2_ARG_CHECK(x && w && y, "%s: null pointer", who);
  B2_ARG_CHECK(b2_conv3x3_halo_dense_supported(N, H, W, C, Cout), "%s: unsupported shape N=%d H=%d W=%d C=%d Cout=%d", who, N,
               H, W, C, Cout);
  B2_ARG_CHECK((col_sum == nullptr) == (col_sumsq == nullptr), "%s: col_sum and col_sumsq go together", who);
  B2_ARG_CHECK(((uintptr_t)x & 15) == 0 && ((uintptr_t)w & 15) == 0 && ((uintptr_t)y & 31) == 0 && ldy % 16 == 0 && ldy >= Cout,
               "%s: x / w 16 B aligned, y 32 B aligned with a row stride that keeps it so", who);
  if (int r = load_encode()) return r;
  HaloGeom g;
  make_geom(&g, N, H, W);
  CUtensorMap tx, tw;
  {
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {64u, (cuuint32_t)g.Wp, (cuuint32_t)(g.TH + 2), 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult cr = g_encode(&tx, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) {
      b2_set_error("%s: cuTensorMapEncodeTiled(x) failed (%d)", who, (int)cr);
      return -3;
    }
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)(kTaps * C), (cuuint64_t)Cout};
    cuuint64_t strides[1] = {(cuuint64_t)(kTaps * C) * 2};
    cuuint32_t box[2] = {64u, (cuuint32_t)Cout};
    cuuint32_t estr[2] = {1, 1};
    CUresult cr = g_encode(&tw, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) {
      b2_set_error("%s: cuTensorMapEncodeTiled(w) failed (%d)", who, (int)cr);
      return -3;
    }
  }
  const int smem = halo_smem_bytes(g, 2, 32);
  static B2PerDeviceMax attr_smem;
  if (attr_smem.below(smem)) {
    B2_CUDA_CHECK(cudaFuncSetAttribute(conv3x3_halo_kernel<false, 2, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_smem.set(smem);
  }
  InBn at = {};
  OutFin fin = {};
  const int grid = g.num_tiles < b2_num_sms() ? g.num_tiles : b2_num_sms();
  conv3x3_halo_kernel<false, 2, 32><<<grid, 64 + 32 * kEpiWarps, smem, (cudaStream_t)stream>>>(tx, tw, (bf16*)y, ldy, g, at,
                                                                                              col_sum, col_sumsq, fin);
  B2_LAUNCH_CHECK("conv3x3_halo_kernel<dense>");
  return 0;
}
