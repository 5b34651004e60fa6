// Direct tcgen05 convolution for the ResNet stem (Conv2d 3->64, 7x7, stride 2, pad 3) -- no patch
// matrix in HBM.
//
// The frames are first repacked (b2_stem_pack) to zero-padded bf16 "units" of 2 pixels x 4 channels
// (16 B; channel 3 is zero), with even and odd padded rows in separate planes:
//     xp[n][plane = y' & 1][hr = y' >> 1][wu = x' >> 1]      y' = y + 3, x' = x + 3
// Output pixel (p, q) gets the flat index u = p * Wq + q (Wq = units per plane row; positions with
// q >= Q are computed and thrown away, ~5 %).  Filter row r of that pixel needs the 8 pixels
// x' = 2q .. 2q+7 of padded row 2p + r, i.e. units  [u + (r >> 1) * Wq, +4)  of plane r & 1:
// the A operand of the implicit GEMM is a TOEPLITZ matrix over the flat unit stream, A[u][k] =
// stream[8u + k].  That is exactly a K-major no-swizzle UMMA shared-memory layout whose core-matrix
// rows are 16 B apart and whose K-direction core-matrix stride (LBO) is ALSO 16 B: overlapping core
// matrices, described to the tensor core purely through the descriptor strides.  One tile of 128
// consecutive u therefore needs only two contiguous spans of the unit stream in shared memory
// (even plane: 3*Wq + 131 units, odd plane: 2*Wq + 131 units; ~9 KB at 112x112), fetched with two
// 1-D bulk copies, and 14 tcgen05.mma (7 filter rows x 2 K=16 halves, K_total = 224 of which 147 are
// real taps) against the 28 KB weight matrix that stays resident in shared memory.
//
// Epilogue: TMEM -> bf16 NHWC rows (64 channels = 128 B contiguous per pixel) + per-channel
// sum / sum-of-squares for the train-mode BatchNorm that follows.
// Reference call site: torchvision ResNet.conv1 inside `self.cnn_backbone(x)`,
// medsos_lrcn/src/models.py:192, lrcn/ucf50-lrcn.py:310.
#include "tc_ptx.cuh"

namespace {
using namespace tc;

constexpr int kCout = 64;
constexpr int kKChunks = 28;                 // 7 filter rows x 4 units of 8 bf16
constexpr int kWBytes = kKChunks * kCout * 16;   // 28672: [chunk][cout][8 bf16], LBO = 1024, SBO = 128
constexpr int kStages = 4;
constexpr int kEpiWarps = 8;
constexpr int kThreads = 64 + 32 * kEpiWarps;
constexpr int kStgBytes = 32 * 64;           // per-warp staging tile: 32 rows x 32 bf16, 64 B rows, XOR-swizzled
constexpr int kSlackUnits = 256;

struct StemGeom {
  int N, H, W, P, Q, Wq, Hq;
  int tiles_per_img;
  int even_bytes, odd_bytes, stage_bytes;
  long img_units;        // 2 * Hq * Wq
  uint32_t magic_wq, magic_tpi;   // ceil(2^32 / d): u / d == __umulhi(u, magic) for u < 2^32 / d (checked on the host)
};

__host__ __device__ inline int align_up(int v, int a) { return (v + a - 1) / a * a; }

StemGeom make_geom(int N, int H, int W) {
  StemGeom g;
  g.N = N; g.H = H; g.W = W;
  g.P = (H + 6 - 7) / 2 + 1;
  g.Q = (W + 6 - 7) / 2 + 1;
  g.Wq = g.Q + 3;
  g.Hq = g.P + 3;
  g.tiles_per_img = (g.P * g.Wq + 127) / 128;
  g.even_bytes = (3 * g.Wq + 131) * 16;
  g.odd_bytes = (2 * g.Wq + 131) * 16;
  g.stage_bytes = align_up(g.even_bytes, 128) + align_up(g.odd_bytes, 128);
  g.img_units = 2L * g.Hq * g.Wq;
  g.magic_wq = (uint32_t)(((1ull << 32) + g.Wq - 1) / g.Wq);
  g.magic_tpi = (uint32_t)(((1ull << 32) + g.tiles_per_img - 1) / g.tiles_per_img);
  return g;
}

template <typename InT>
__global__ void __launch_bounds__(256)
stem_pack_kernel(const InT* __restrict__ x, uint4* __restrict__ xp, StemGeom g) {
  const long total = g.N * g.img_units;
  const long plane_sz = (long)g.H * g.W;
  for (long v = (long)blockIdx.x * blockDim.x + threadIdx.x; v < total + kSlackUnits;
       v += (long)gridDim.x * blockDim.x) {
    uint4 o = make_uint4(0u, 0u, 0u, 0u);
    if (v < total) {
      const int wu = (int)(v % g.Wq);
      long t = v / g.Wq;
      const int hr = (int)(t % g.Hq);
      t /= g.Hq;
      const int plane = (int)(t & 1);
      const long n = t >> 1;
      const int y = 2 * hr + plane - 3;
      const int x0 = 2 * wu - 3;
      if (y >= 0 && y < g.H) {
        const InT* row = x + n * 3 * plane_sz + (long)y * g.W;
        float a[2][3];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int xx = x0 + j;
          const bool ok = xx >= 0 && xx < g.W;
#pragma unroll
          for (int c = 0; c < 3; ++c) a[j][c] = ok ? (float)row[c * plane_sz + xx] : 0.f;
        }
        o.x = pack_bf16x2(a[0][0], a[0][1]);
        o.y = pack_bf16x2(a[0][2], 0.f);
        o.z = pack_bf16x2(a[1][0], a[1][1]);
        o.w = pack_bf16x2(a[1][2], 0.f);
      }
    }
    xp[v] = o;
  }
}

__global__ void __launch_bounds__(kThreads, 1)
stem_conv_kernel(const uint4* __restrict__ xp, const uint4* __restrict__ wk, bf16* __restrict__ y, StemGeom g,
                 float* col_sum, float* col_sumsq) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
  uint8_t* w_s = smem;                                   // resident weights
  uint8_t* stage_s = smem + kWBytes;                     // kStages x stage_bytes
  uint8_t* after = stage_s + kStages * g.stage_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(after);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tfull_bar = empty_bar + kStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* w_bar = tempty_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_bar + 1);
  float* stat_s = reinterpret_cast<float*>(after + 128);            // [4 quarters x 2 half-warps][2][64]
  uint8_t* staging_s = after + 128 + 8 * 2 * kCout * 4;            // kEpiWarps x kStgBytes
  const bool want_stats = col_sum != nullptr;
  for (int i = threadIdx.x; i < 16 * kCout; i += kThreads) stat_s[i] = 0.f;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = g.N * g.tiles_per_img;
  const int even_span = align_up(g.even_bytes, 128);

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], kEpiWarps);
    }
    mbar_init(w_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tc_alloc(tmem_slot, 2 * kCout);
    tc_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // =========================== bulk-copy producer ===========================
    if (lane == 0) {
      mbar_expect_tx(w_bar, kWBytes);
      bulk_load_1d(w_s, wk, kWBytes, w_bar);
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int img = g.tiles_per_img == 1 ? tile : (int)__umulhi((uint32_t)tile, g.magic_tpi);
        const int u0 = (tile - img * g.tiles_per_img) * 128;
        const uint4* src_even = xp + (long)img * g.img_units + u0;
        const uint4* src_odd = src_even + (long)g.Hq * g.Wq;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* st = stage_s + stage * g.stage_bytes;
        mbar_expect_tx(&full_bar[stage], (uint32_t)(g.even_bytes + g.odd_bytes));
        bulk_load_1d(st, src_even, (uint32_t)g.even_bytes, &full_bar[stage]);
        bulk_load_1d(st + even_span, src_odd, (uint32_t)g.odd_bytes, &full_bar[stage]);
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    constexpr uint32_t idesc = make_idesc(128, kCout);
    mbar_wait(w_bar, 0);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    const uint32_t w_addr = smem_u32(w_s);
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * kCout);
        const uint32_t st = smem_u32(stage_s + stage * g.stage_bytes);
#pragma unroll
        for (int r = 0; r < 7; ++r) {
          const uint32_t a_row = st + ((r & 1) ? (uint32_t)even_span : 0u) + (uint32_t)((r >> 1) * g.Wq * 16);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const uint64_t da = make_nosw_desc(a_row + h * 32, 16, 128);
            const uint64_t db = make_nosw_desc(w_addr + (uint32_t)((r * 4 + h * 2) * kCout * 16), kCout * 16, 128);
            tc_mma_bf16(d_tmem, da, db, idesc, (r | h) != 0);
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
  } else {
    // =========================== epilogue (warps 2..9) ===========================
    const int quarter = warp & 3;
    const int ch = (warp - 2) >> 2;        // which 32-channel half this warp stores
    uint8_t* stg = staging_s + (warp - 2) * kStgBytes;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int img = g.tiles_per_img == 1 ? tile : (int)__umulhi((uint32_t)tile, g.magic_tpi);
      const int u = (tile - img * g.tiles_per_img) * 128 + quarter * 32 + lane;
      const int p = (int)__umulhi((uint32_t)u, g.magic_wq);
      const int q = u - p * g.Wq;
      const bool row_ok = p < g.P && q < g.Q;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      uint32_t raw[32];
      tc_ld32(tmem_base + (uint32_t)(acc * kCout + ch * 32) + ((uint32_t)(quarter * 32) << 16), raw);
      tc_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);    // accumulator is in registers: release it early
      uint32_t pk[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) pk[j] = pack_bf16x2(__uint_as_float(raw[2 * j]), __uint_as_float(raw[2 * j + 1]));
      if (row_ok) {
        bf16* dp = y + ((((long)img * g.P + p) * g.Q + q) * kCout + ch * 32);
        asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dp), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]),
                     "r"(pk[3]), "r"(pk[4]), "r"(pk[5]), "r"(pk[6]), "r"(pk[7])
                     : "memory");
        asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dp + 16), "r"(pk[8]), "r"(pk[9]),
                     "r"(pk[10]), "r"(pk[11]), "r"(pk[12]), "r"(pk[13]), "r"(pk[14]), "r"(pk[15])
                     : "memory");
      }
      if (want_stats) {
        if (!row_ok) {
#pragma unroll
          for (int j = 0; j < 16; ++j) pk[j] = 0u;
        }
        const int sw = (lane >> 1) & 3;
#pragma unroll
        for (int c = 0; c < 4; ++c)
          *reinterpret_cast<uint4*>(stg + lane * 64 + ((c ^ sw) << 4)) =
              make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
        __syncwarp();
        // lane (w = lane & 15, hf = lane >> 4) sums columns 2w, 2w+1 over rows 16 hf .. 16 hf + 15 straight from
        // the swizzled tile (same scheme as gemm_tc.cu: packed fp32 adds, the two half-warps never share a bank)
        const int sw_w = lane & 15, sw_hf = lane >> 4;
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
        // each half-warp owns its own copy of the statistics (no shuffle, no atomics: one writer per slot)
        float2* st = reinterpret_cast<float2*>(stat_s + (quarter * 2 + sw_hf) * 2 * kCout + ch * 32 + 2 * sw_w);
        float2* st2 = reinterpret_cast<float2*>(stat_s + (quarter * 2 + sw_hf) * 2 * kCout + kCout + ch * 32 + 2 * sw_w);
        *st = __fadd2_rn(*st, __fadd2_rn(s1a, s1b));
        *st2 = __fadd2_rn(*st2, __fadd2_rn(s2a, s2b));
        __syncwarp();
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (want_stats) {
    for (int c = threadIdx.x; c < kCout; c += kThreads) {
      float a1 = 0.f, a2 = 0.f;
#pragma unroll
      for (int qq = 0; qq < 8; ++qq) {
        a1 += stat_s[qq * 2 * kCout + c];
        a2 += stat_s[qq * 2 * kCout + kCout + c];
      }
      atomicAdd(col_sum + c, a1);
      atomicAdd(col_sumsq + c, a2);
    }
  }
  if (warp == 1) {
    tc_fence_after();
    tc_dealloc(tmem_base, 2 * kCout);
  }
}

int stem_smem_bytes(const StemGeom& g) {
  return kWBytes + kStages * g.stage_bytes + 128 + 8 * 2 * kCout * 4 + kEpiWarps * kStgBytes + 256;
}

}  // namespace

B2_API long b2_stem_packed_bytes(int N, int H, int W) {
  if (N <= 0 || H <= 0 || W <= 0) return 0;
  const StemGeom g = make_geom(N, H, W);
  return (g.N * g.img_units + kSlackUnits) * 16;
}

B2_API int b2_stem_pack(const void* x, int in_bf16, void* xp, int N, int H, int W, void* stream) {
  B2_ARG_CHECK(x && xp && N > 0 && H > 0 && W > 0, "b2_stem_pack: null pointer or empty");
  const StemGeom g = make_geom(N, H, W);
  const long total = g.N * g.img_units + kSlackUnits;
  long blocks = (total + 255) / 256;
  const long cap = (long)b2_num_sms() * 16;
  const int grid = (int)(blocks < cap ? blocks : cap);
  if (in_bf16)
    stem_pack_kernel<bf16><<<grid, 256, 0, (cudaStream_t)stream>>>((const bf16*)x, (uint4*)xp, g);
  else
    stem_pack_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)x, (uint4*)xp, g);
  B2_LAUNCH_CHECK("stem_pack_kernel");
  return 0;
}

B2_API int b2_stem_conv_bf16(const void* xp, const void* wk, void* y, int N, int H, int W, float* col_sum,
                             float* col_sumsq, void* stream) {
  B2_ARG_CHECK(xp && wk && y && N > 0 && H > 0 && W > 0, "b2_stem_conv_bf16: null pointer or empty");
  B2_ARG_CHECK((col_sum == nullptr) == (col_sumsq == nullptr), "b2_stem_conv_bf16: col_sum and col_sumsq go together");
  B2_ARG_CHECK(((uintptr_t)xp & 15) == 0 && ((uintptr_t)wk & 15) == 0 && ((uintptr_t)y & 31) == 0,
               "b2_stem_conv_bf16: xp/wk must be 16 B and y 32 B aligned");
  const StemGeom g = make_geom(N, H, W);
  const int smem = stem_smem_bytes(g);
  B2_ARG_CHECK(smem <= 227 * 1024, "b2_stem_conv_bf16: frame width %d too large for the shared-memory stage", W);
  B2_ARG_CHECK((long)g.N * g.tiles_per_img < (1L << 31), "b2_stem_conv_bf16: too many tiles");
  B2_ARG_CHECK((unsigned long)g.N * g.tiles_per_img * g.tiles_per_img < (1ul << 32) &&
                   (unsigned long)(g.tiles_per_img * 128 + 128) * g.Wq < (1ul << 32),
               "b2_stem_conv_bf16: shape outside the multiply-high division range");
  static B2PerDeviceMax attr_smem;
  if (attr_smem.below((int)smem)) {
    B2_CUDA_CHECK(cudaFuncSetAttribute(stem_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_smem.set((int)smem);
  }
  const int tiles = g.N * g.tiles_per_img;
  const int grid = tiles < b2_num_sms() ? tiles : b2_num_sms();
  stem_conv_kernel<<<grid, kThreads, smem, (cudaStream_t)stream>>>((const uint4*)xp, (const uint4*)wk, (bf16*)y, g,
                                                                  col_sum, col_sumsq);
  B2_LAUNCH_CHECK("stem_conv_kernel");
  return 0;
}
