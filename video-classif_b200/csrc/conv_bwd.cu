// Backward kernels of the trainable frame encoder (full / partial fine-tune of the torchvision ResNet:
// lrcn/rgb_lrcn.py:208-245, lrcn/lrcn.py:246-283 `freeze_cnn_layers`, medsos models.py:144-145 when un-frozen).
//
//   * conv_wgrad_kernel     dW[co, r, s, ci] = sum_m dy[m, co] * x[pix(m) + (r, s), ci]  on the tensor core with BOTH
//                           operands MN-major: the [64 pixels x 64 channels] SWIZZLE_128B tiles that TMA delivers for
//                           the forward (channels contiguous) are exactly the transposed operands the weight gradient
//                           needs, so neither dy^T nor an im2col^T matrix is ever materialised.  The reduction
//                           dimension is the pixel index; the filter tap (r, s) is a shift of the im2col-mode TMA
//                           box.  One CTA = (tap, 128 output channels, up to 256 input channels, pixel split);
//                           the fp32 accumulator lives in TMEM over all of the CTA's pixel tiles and is drained once
//                           with vector reductions into dW.
//   * bn_bwd_reduce_kernel  ReLU mask (dzm = dz * [z > 0]) + per-channel sum(dzm), sum(dzm * xhat)
//   * bn_bwd_apply_kernel   dy = gamma * invstd * (dz - mean(dz) - xhat * mean(dz * xhat))      (train-mode BatchNorm)
//   * dilate2 / maxpool / avgpool backward helpers
// The data gradient of a convolution is a convolution of dy with the transposed, flipped filter: it runs on the forward
// kernels of gemm_tc.cu (stride 2: over the zero-dilated dy).
#include "tc_ptx.cuh"

#include <cudaTypedefs.h>
#include <mutex>

namespace {
using namespace tc;

constexpr int kRows = 64;                   // pixels (reduction rows) per pipeline stage
constexpr int kPanel = kRows * 128;         // one [64 pixels x 64 channels] bf16 panel, SWIZZLE_128B: 8 KB
constexpr int kWgThreads = 64 + 128;        // TMA warp, MMA warp, 4 drain warps

// MT = 128-row output-channel tiles per CTA: MT = 2 (two accumulators sharing the x panels, 512 TMEM columns) raises the
// operand intensity 1.5x for the wide layers (the kernel is bound by L2 -> SM operand delivery)
template <int BNC, int MT = 1>
struct WgLayout {
  static constexpr int kBPanels = BNC / 64;
  static constexpr int kAPanels = 2 * MT;
  static constexpr int kStageBytes = (kAPanels + kBPanels) * kPanel;   // 24 / 32 / 48 KB (MT = 2: 64 KB)
  static constexpr int kStages = MT == 2 ? 3 : (BNC == 256 ? 4 : (BNC == 128 ? 6 : 8));
  static constexpr int kBarOffset = kStages * kStageBytes;
  static constexpr int kSmem = kBarOffset + (2 * kStages + 1) * 8 + 16 + 1024;
  static_assert(kSmem <= 227 * 1024, "shared memory budget");
};

struct WgGeom {
  int is_im2col;
  int P, Q, S, stride, lower_w, lower_h;
  int taps;
  int Mp;          // output pixels N * P * Q
  int Cout, C;
  int ci_blocks;
  int tpg;         // filter taps per CTA: narrow layers (C <= 128) put 256 / C taps side by side in the N dimension of one
  int ppt;         //   MMA (the dy tile is fetched once for all of them); ppt = 64-channel panels per tap
  int tap_groups;
};

// MN-major SWIZZLE_128B descriptor (same encoding as the Gram kernel of gemm_tc.cu): 64 contiguous MN elements per
// 128 B row (one row per K index), 8-row groups SBO = 1024 B apart, 64-element MN blocks LBO bytes apart.
__device__ __forceinline__ uint64_t sw128_mn_desc(uint32_t smem_addr, uint32_t lbo) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

template <int BNC, int MT = 1>
__global__ void __launch_bounds__(kWgThreads, 1)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap tmap_dy, const __grid_constant__ CUtensorMap tmap_x, WgGeom g,
                  float* __restrict__ dw) {
  using L = WgLayout<BNC, MT>;
  static_assert(MT * BNC <= 512, "TMEM columns");
  constexpr int kStages = L::kStages;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* done_bar = empty_bar + kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done_bar + 1);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tap0 = (blockIdx.x % g.tap_groups) * g.tpg;
  const int split = blockIdx.x / g.tap_groups;
  const int splits = gridDim.x / g.tap_groups;
  const int ci_blk = blockIdx.y % g.ci_blocks;
  const int co_blk = blockIdx.y / g.ci_blocks;
  const int num_tiles = (g.Mp + kRows - 1) / kRows;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_dy);
    prefetch_tmap(&tmap_x);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(done_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tc_alloc(tmem_slot, MT * BNC);
    tc_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int valid = 0;                                   // panels that hold a real (tap, channel slab)
      for (int pp = 0; pp < L::kBPanels; ++pp)
        valid += (tap0 + pp / g.ppt < g.taps) && (ci_blk * BNC + (pp % g.ppt) * 64 < g.C);
      const uint32_t stage_tx = (uint32_t)((L::kAPanels + valid) * kPanel);
      for (int t = split; t < num_tiles; t += splits) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * L::kStageBytes;
        uint8_t* sb = sa + L::kAPanels * kPanel;
        mbar_expect_tx(&full_bar[stage], stage_tx);
        const int m0 = t * kRows;
#pragma unroll
        for (int pa = 0; pa < L::kAPanels; ++pa)                                       // columns past Cout: zero fill
          tma_load_2d(sa + pa * kPanel, &tmap_dy, &full_bar[stage], co_blk * (128 * MT) + pa * 64, m0);
        int cn = 0, cw = 0, ch = 0;
        if (g.is_im2col) {
          const int pq = g.P * g.Q;
          cn = m0 / pq;
          const int rem = m0 - cn * pq;
          const int p = rem / g.Q;
          const int q = rem - p * g.Q;
          cw = g.lower_w + q * g.stride;
          ch = g.lower_h + p * g.stride;
        }
#pragma unroll
        for (int pp = 0; pp < L::kBPanels; ++pp) {
          const int tap = tap0 + pp / g.ppt;
          const int c0 = ci_blk * BNC + (pp % g.ppt) * 64;
          if (tap >= g.taps || c0 >= g.C) continue;          // the MMA reads stale shared memory there: never drained
          if (g.is_im2col) {
            const int r = tap / g.S;
            const int s = tap - r * g.S;
            tma_load_im2col_4d(sb + pp * kPanel, &tmap_x, &full_bar[stage], c0, cw, ch, cn, (uint16_t)s, (uint16_t)r);
          } else {
            tma_load_2d(sb + pp * kPanel, &tmap_x, &full_bar[stage], c0, m0);
          }
        }
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // D[co, ci] += sum over the 64 pixel rows of dy[row, co] * x[row, ci]: both operands MN-major (bits 15, 16)
    constexpr uint32_t idesc = make_idesc(128, BNC) | (1u << 15) | (1u << 16);
    int stage = 0;
    uint32_t phase = 0;
    bool first = true;
    for (int t = split; t < num_tiles; t += splits) {
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t sa = smem_u32(smem + stage * L::kStageBytes);
        const uint32_t sb = sa + L::kAPanels * kPanel;
#pragma unroll
        for (int ks = 0; ks < kRows / 16; ++ks) {            // 16 pixel rows (two 8-row groups) per MMA
          const uint64_t db = sw128_mn_desc(sb + ks * 2048, kPanel);
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            const uint64_t da = sw128_mn_desc(sa + mt * 2 * kPanel + ks * 2048, kPanel);
            tc_mma_bf16(tmem_base + (uint32_t)(mt * BNC), da, db, idesc, !(first && ks == 0));
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
    // drain: lane = output channel row of the accumulator, 32-column chunks of input channels -> dW[co][tap][ci]
    const bool any = split < num_tiles;
    if (any) {
      mbar_wait(done_bar, 0);
      tc_fence_after();
      const int quarter = warp & 3;
#pragma unroll 1
      for (int mt = 0; mt < MT; ++mt) {
      const int co = co_blk * (128 * MT) + mt * 128 + quarter * 32 + lane;
#pragma unroll 1
      for (int ch = 0; ch < BNC / 32; ++ch) {
        const int pp = ch >> 1;
        const int tap = tap0 + pp / g.ppt;
        const int ci0 = ci_blk * BNC + (pp % g.ppt) * 64 + (ch & 1) * 32;
        if (tap >= g.taps || ci0 >= g.C) continue;     // warp-uniform
        float* dst_row = dw + ((long)co * g.taps + tap) * g.C;
        uint32_t raw[32];
        tc_ld32(tmem_base + (uint32_t)(mt * BNC + ch * 32) + ((uint32_t)(quarter * 32) << 16), raw);
        tc_wait_ld();
        if (co < g.Cout) {
          if (ci0 + 32 <= g.C) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst_row + ci0 + j),
                           "f"(__uint_as_float(raw[j])), "f"(__uint_as_float(raw[j + 1])), "f"(__uint_as_float(raw[j + 2])),
                           "f"(__uint_as_float(raw[j + 3]))
                           : "memory");
          } else {
            for (int j = 0; j < 32; ++j)
              if (ci0 + j < g.C) atomicAdd(dst_row + ci0 + j, __uint_as_float(raw[j]));
          }
        }
      }
      }
      tc_fence_before();
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tc_dealloc(tmem_base, MT * BNC);
  }
}

// ------------------------------------------------------------------------------------------------ BatchNorm backward
// [M, C] bf16, C % 8 == 0, C <= 2048.  Thread = 8 consecutive channels (one 16-byte load), fixed over all its rows.
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  f[0] = __uint_as_float(u.x << 16); f[1] = __uint_as_float(u.x & 0xffff0000u);
  f[2] = __uint_as_float(u.y << 16); f[3] = __uint_as_float(u.y & 0xffff0000u);
  f[4] = __uint_as_float(u.z << 16); f[5] = __uint_as_float(u.z & 0xffff0000u);
  f[6] = __uint_as_float(u.w << 16); f[7] = __uint_as_float(u.w & 0xffff0000u);
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}

// mean / invstd of the forward: from (sum, sumsq) of the raw conv output over `count` rows (train) or the running
// statistics (eval)
__device__ __forceinline__ void bn_mean_invstd(const float* sum, const float* sumsq, const float* rmean, const float* rvar,
                                               int c, float inv_count, float eps, int train, float& mean, float& invstd) {
  if (train) {
    mean = sum[c] * inv_count;
    const float var = fmaxf(sumsq[c] * inv_count - mean * mean, 0.f);
    invstd = rsqrtf(var + eps);
  } else {
    mean = rmean[c];
    invstd = rsqrtf(rvar[c] + eps);
  }
}

__global__ void __launch_bounds__(256)
bn_bwd_reduce_kernel(const bf16* dz, bf16* dzm, const bf16* __restrict__ z, const bf16* __restrict__ y,
                     const float* __restrict__ sum, const float* __restrict__ sumsq, const float* __restrict__ rmean,
                     const float* __restrict__ rvar, float* __restrict__ out_s1, float* __restrict__ out_s2, long M, int C,
                     int rows_per_block, float inv_count, float eps, int train, float hi) {
  __shared__ float red[2][256 * 8];                 // [s1 | s2][thread][8 channels]: 16 KB
  const int groups = C >> 3;                         // channel groups per row
  const int lanes = 256 / groups > 0 ? 256 / groups : 1;   // rows handled concurrently by the block
  const int grp = threadIdx.x % groups;
  const int rl = threadIdx.x / groups;
  const bool active = rl < lanes && groups <= 256;
  float mean[8], invstd[8], s1[8], s2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    s1[j] = s2[j] = 0.f;
    if (active) bn_mean_invstd(sum, sumsq, rmean, rvar, grp * 8 + j, inv_count, eps, train, mean[j], invstd[j]);
  }
  const long r0 = (long)blockIdx.x * rows_per_block;
  const long r1 = min(M, r0 + rows_per_block);
  if (active) {
    for (long r = r0 + rl; r < r1; r += lanes) {
      const long off = r * C + grp * 8;
      uint4 g = *reinterpret_cast<const uint4*>(dz + off);
      const uint4 yv = *reinterpret_cast<const uint4*>(y + off);
      float gf[8], yf[8];
      unpack8(g, gf);
      unpack8(yv, yf);
      if (z != nullptr) {
        const uint4 zv = *reinterpret_cast<const uint4*>(z + off);
        float zf[8];
        unpack8(zv, zf);
#pragma unroll
        for (int j = 0; j < 8; ++j) gf[j] = (zf[j] > 0.f && zf[j] < hi) ? gf[j] : 0.f;     // ReLU (hi = inf) / ReLU6 (hi = 6)
        g.x = pack2(gf[0], gf[1]); g.y = pack2(gf[2], gf[3]); g.z = pack2(gf[4], gf[5]); g.w = pack2(gf[6], gf[7]);
        if (dzm != nullptr) *reinterpret_cast<uint4*>(dzm + off) = g;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s1[j] += gf[j];
        s2[j] = fmaf(gf[j], (yf[j] - mean[j]) * invstd[j], s2[j]);
      }
    }
  }
  // block reduction over the row lanes, then one atomic per channel
  float* r_s1 = &red[0][0];
  float* r_s2 = &red[1][0];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    r_s1[threadIdx.x * 8 + j] = active ? s1[j] : 0.f;
    r_s2[threadIdx.x * 8 + j] = active ? s2[j] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += 256) {
    const int gq = i >> 3, j = i & 7;
    float a = 0.f, b = 0.f;
    for (int l = 0; l < lanes; ++l) {
      a += r_s1[(l * groups + gq) * 8 + j];
      b += r_s2[(l * groups + gq) * 8 + j];
    }
    atomicAdd(out_s1 + i, a);
    atomicAdd(out_s2 + i, b);
  }
}

// dy = A[c] * dz + B[c] * y + K[c]  with  A = gamma * invstd,  B = -A * invstd * s2 / count,  K = -A * s1 / count - B * mean
// (train) or dy = A[c] * dz (eval).  A thread's 8 channels are the same for every row it visits (the grid stride is a
// multiple of the channel-group count), so the coefficients are computed once.  z != NULL: dz is masked by (z > 0) here.
__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(const bf16* __restrict__ dz, const bf16* __restrict__ z, const bf16* __restrict__ y,
                    bf16* __restrict__ dy, const float* __restrict__ gamma, const float* __restrict__ sum,
                    const float* __restrict__ sumsq, const float* __restrict__ rmean, const float* __restrict__ rvar,
                    const float* __restrict__ s1, const float* __restrict__ s2, long M, int C, float inv_count, float eps,
                    int train, float hi) {
  const int groups = C >> 3;
  const long total = M * groups;
  const long i0 = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int grp = (int)(i0 % groups);
  float ca[8], cb[8], ck[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = grp * 8 + j;
    float mean, invstd;
    bn_mean_invstd(sum, sumsq, rmean, rvar, c, inv_count, eps, train, mean, invstd);
    ca[j] = gamma[c] * invstd;
    cb[j] = train ? -ca[j] * invstd * s2[c] * inv_count : 0.f;
    ck[j] = train ? -ca[j] * s1[c] * inv_count - cb[j] * mean : 0.f;
  }
  for (long i = i0; i < total; i += (long)gridDim.x * blockDim.x) {
    const long off = (i / groups) * C + grp * 8;
    float gf[8], yf[8], o[8];
    unpack8(*reinterpret_cast<const uint4*>(dz + off), gf);
    unpack8(*reinterpret_cast<const uint4*>(y + off), yf);
    if (z != nullptr) {
      float zf[8];
      unpack8(*reinterpret_cast<const uint4*>(z + off), zf);
#pragma unroll
      for (int j = 0; j < 8; ++j) gf[j] = (zf[j] > 0.f && zf[j] < hi) ? gf[j] : 0.f;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = fmaf(ca[j], gf[j], fmaf(cb[j], yf[j], ck[j]));
    uint4 r;
    r.x = pack2(o[0], o[1]); r.y = pack2(o[2], o[3]); r.z = pack2(o[4], o[5]); r.w = pack2(o[6], o[7]);
    *reinterpret_cast<uint4*>(dy + off) = r;
  }
}

// z[n, 2p, 2q, :] = dy[n, p, q, :] (zero elsewhere), z is [N, H, W, C] (pre-zeroed by the caller)
__global__ void __launch_bounds__(256)
dilate2_kernel(const bf16* __restrict__ dy, bf16* __restrict__ z, int N, int P, int Q, int H, int W, int C) {
  const int groups = C >> 3;
  const long total = (long)N * P * Q * groups;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int grp = (int)(i % groups);
    long px = i / groups;
    const int q = (int)(px % Q);
    px /= Q;
    const int p = (int)(px % P);
    const long n = px / P;
    if (2 * p < H && 2 * q < W)
      *reinterpret_cast<uint4*>(z + (((n * H + 2 * p) * W + 2 * q) * C + grp * 8)) =
          *reinterpret_cast<const uint4*>(dy + (i / groups) * C + grp * 8);
  }
}

// Both kernel layouts of one torch conv weight w [Cout, Cin, R, S] fp32 in a single pass:
//   wk [Cout, R, S, Cin]        bf16  the forward (and weight-gradient) layout: K index = (r S + s) Cin + ci
//   wt [Cin, R, S, Cout]        bf16  the data-gradient layout: flipped taps, channels swapped (wt[ci][r][s][co] = w[co][ci][R-1-r][S-1-s])
// (replaces permute / flip / contiguous / cast chains of ~5 ATen kernels per convolution and step)
__global__ void __launch_bounds__(256)
conv_weight_layouts_kernel(const float* __restrict__ w, bf16* __restrict__ wk, bf16* __restrict__ wt, int Cout, int Cin, int R,
                           int S) {
  const long total = (long)Cout * Cin * R * S;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int s = (int)(i % S);
    long t = i / S;
    const int r = (int)(t % R);
    t /= R;
    const int ci = (int)(t % Cin);
    const long co = t / Cin;
    const bf16 v = __float2bfloat16(w[i]);
    if (wk != nullptr) wk[((co * R + r) * S + s) * Cin + ci] = v;
    if (wt != nullptr) wt[(((long)ci * R + (R - 1 - r)) * S + (S - 1 - s)) * Cout + co] = v;
  }
}

// dz[n, hw, c] = dfeat[n, c] / HW  (backward of the global average pool; fp32 -> bf16)
__global__ void __launch_bounds__(256)
avgpool_bwd_kernel(const float* __restrict__ dfeat, bf16* __restrict__ dz, long N, int HW, int C) {
  const long total = N * HW * C;
  const float inv = 1.f / (float)HW;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long n = i / ((long)HW * C);
    dz[i] = __float2bfloat16(dfeat[n * C + c] * inv);
  }
}

// Backward of relu(bn(raw)) -> maxpool 3x3 / stride 2 / pad 1 (the stem tail): for every pooled output the gradient goes
// to the FIRST maximal element of its window (torch's rule) of a = relu(raw * scale + shift); draw[n,h,w,c] accumulates
// (fp32, pre-zeroed).  The ReLU mask (a > 0) is applied here, so `draw` is the gradient w.r.t. the BatchNorm output.
__global__ void __launch_bounds__(256)
maxpool_relu_bwd_kernel(const bf16* __restrict__ raw, const float* __restrict__ scale, const float* __restrict__ shift,
                        const bf16* __restrict__ dpool, float* __restrict__ dbn, int N, int H, int W, int P, int Q, int C) {
  const long total = (long)N * P * Q * C;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long px = i / C;
    const int q = (int)(px % Q);
    px /= Q;
    const int p = (int)(px % P);
    const long n = px / P;
    const float sc = scale[c], sh = shift[c];
    // relu(raw * sc + sh) is monotone in raw (rising for sc >= 0, falling otherwise): the first maximal element of the
    // window is the first extremal RAW value -- exact bf16 comparisons, no dependence on how the forward rounded
    const float sgn = sc >= 0.f ? 1.f : -1.f;
    float best = -INFINITY;
    long best_off = -1;
    for (int dh = 0; dh < 3; ++dh) {
      const int h = 2 * p - 1 + dh;
      if (h < 0 || h >= H) continue;
      for (int dw_ = 0; dw_ < 3; ++dw_) {
        const int w = 2 * q - 1 + dw_;
        if (w < 0 || w >= W) continue;
        const long off = ((n * H + h) * W + w) * C + c;
        const float key = sgn * __bfloat162float(raw[off]);
        if (key > best) {
          best = key;
          best_off = off;
        }
      }
    }
    if (best_off < 0) continue;
    best = fmaf(__bfloat162float(raw[best_off]), sc, sh);
    if (best > 0.f) atomicAdd(dbn + best_off, __bfloat162float(dpool[i]));
  }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t, const cuuint32_t*,
                                   CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                   CUtensorMapFloatOOBfill);
EncodeTiledFn g_tiled = nullptr;
EncodeIm2colFn g_im2col = nullptr;
std::once_flag g_once;

int load_encoders() {
  std::call_once(g_once, [] {
    cudaDriverEntryPointQueryResult qres;
    void* fn = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      g_tiled = reinterpret_cast<EncodeTiledFn>(fn);
    fn = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      g_im2col = reinterpret_cast<EncodeIm2colFn>(fn);
  });
  if (g_tiled == nullptr || g_im2col == nullptr) {
    b2_set_error("cuTensorMapEncode* driver entry points unavailable (no CUDA driver / GPU?)");
    return -2;
  }
  return 0;
}

int tmap_rows64(CUtensorMap* map, const void* base, long rows, long cols, long ld) {
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)kRows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_tiled(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    b2_set_error("conv_bwd: cuTensorMapEncodeTiled failed (%d): rows=%ld cols=%ld ld=%ld", (int)r, rows, cols, ld);
    return -3;
  }
  return 0;
}

template <int BNC, int MT = 1>
int launch_wgrad(const CUtensorMap& tdy, const CUtensorMap& tx, const WgGeom& g, float* dw, cudaStream_t stream) {
  using L = WgLayout<BNC, MT>;
  static B2PerDeviceOnce attr_set;
  if (attr_set.needed()) {
    B2_CUDA_CHECK(cudaFuncSetAttribute(conv_wgrad_kernel<BNC, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kSmem));
    attr_set.mark();
  }
  const int co_blocks = b2_ceil_div(g.Cout, 128 * MT);
  const int units = g.tap_groups * co_blocks * g.ci_blocks;
  const int tiles = b2_ceil_div(g.Mp, kRows);
  int splits = b2_num_sms() / units;          // one CTA per SM and a single wave
  if (splits > tiles) splits = tiles;
  if (splits < 1) splits = 1;
  dim3 grid((unsigned)(g.tap_groups * splits), (unsigned)(co_blocks * g.ci_blocks));
  conv_wgrad_kernel<BNC, MT><<<grid, kWgThreads, L::kSmem, stream>>>(tdy, tx, g, dw);
  B2_LAUNCH_CHECK("conv_wgrad_kernel");
  return 0;
}

unsigned ew_blocks(long n) {
  long b = (n + 255) / 256;
  const long cap = (long)b2_num_sms() * 16;
  return (unsigned)(b < cap ? (b > 0 ? b : 1) : cap);
}

}  // namespace

// dw[Cout, R, S, C] (fp32, ACCUMULATED: the caller zeroes it) += sum over output pixels of dy[m, co] * x[pix(m)+(r,s), ci]
B2_API int b2_conv2d_wgrad_nhwc_bf16(const void* x, int Nimg, int H, int W, int C, const void* dy, int Cout, int R, int S,
                                     int stride, int pad, float* dw, void* stream) {
  const char* who = "b2_conv2d_wgrad_nhwc_bf16";
  B2_ARG_CHECK(x && dy && dw && Nimg > 0 && H > 0 && W > 0 && C > 0 && Cout > 0, "%s: null pointer or empty shape", who);
  B2_ARG_CHECK(C % 8 == 0 && Cout % 8 == 0, "%s: C and Cout must be multiples of 8 (16-byte rows)", who);
  B2_ARG_CHECK(((uintptr_t)x & 15) == 0 && ((uintptr_t)dy & 15) == 0 && ((uintptr_t)dw & 15) == 0, "%s: 16 B alignment", who);
  B2_ARG_CHECK(R >= 1 && S >= 1 && R <= 7 && S <= 7 && stride >= 1 && stride <= 8 && pad >= 0 && pad <= 3,
               "%s: unsupported filter geometry R=%d S=%d stride=%d pad=%d", who, R, S, stride, pad);
  const int P = (H + 2 * pad - R) / stride + 1;
  const int Q = (W + 2 * pad - S) / stride + 1;
  B2_ARG_CHECK(P > 0 && Q > 0, "%s: empty output", who);
  const long Ml = (long)Nimg * P * Q;
  B2_ARG_CHECK(Ml < (1L << 31), "%s: too many output pixels", who);
  if (int r = load_encoders()) return r;
  const bool plain = R == 1 && S == 1 && stride == 1 && pad == 0;
  B2_ARG_CHECK(plain || C % 64 == 0, "%s: a filter window needs C %% 64 == 0 (got %d)", who, C);
  CUtensorMap tdy, tx;
  if (int r = tmap_rows64(&tdy, dy, Ml, Cout, Cout)) return r;
  if (plain) {
    if (int r = tmap_rows64(&tx, x, Ml, C, C)) return r;
  } else {
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)Nimg};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    int lower[2] = {-pad, -pad};
    int upper[2] = {pad - (S - 1), pad - (R - 1)};
    cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
    CUresult cr = g_im2col(&tx, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), dims, strides, lower, upper, 64u,
                           (cuuint32_t)kRows, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) {
      b2_set_error("%s: cuTensorMapEncodeIm2col failed (%d): N=%d H=%d W=%d C=%d R=%d S=%d stride=%d pad=%d", who, (int)cr,
                   Nimg, H, W, C, R, S, stride, pad);
      return -3;
    }
    int drv = 0;          // same driver quirk as the forward im2col maps (gemm_tc.cu)
    cudaDriverGetVersion(&drv);
    if (drv <= 13010 && (long)Nimg * H * W * C * 2 < 131072) reinterpret_cast<uint64_t*>(&tx)[1] &= ~(1ull << 21);
  }
  const int taps = R * S;
  const bool grouped = taps > 1 && C <= 128;         // C = 64: 4 taps per CTA, C = 128: 2 taps per CTA (N = 256)
  const int bnc = grouped ? 256 : (C > 128 ? 256 : (C > 64 ? 128 : 64));
  const int tpg = grouped ? 256 / C : 1;
  WgGeom g = {plain ? 0 : 1, P, Q, S, stride, -pad, -pad, taps, (int)Ml, Cout, C, grouped ? 1 : b2_ceil_div(C, bnc),
              tpg,  grouped ? C / 64 : bnc / 64, b2_ceil_div(taps, tpg)};
  cudaStream_t st = (cudaStream_t)stream;
  // wide layers: 256 output channels per CTA (two accumulators share the x panels) while enough work units remain to
  // fill the machine with pixel splits
  static const bool no_mt2 = getenv("B2_WGRAD_NO_MT2") != nullptr;
  if (bnc == 256 && Cout >= 256 && !no_mt2) return launch_wgrad<256, 2>(tdy, tx, g, dw, st);
  switch (bnc) {
    case 256: return launch_wgrad<256>(tdy, tx, g, dw, st);
    case 128: return launch_wgrad<128>(tdy, tx, g, dw, st);
    default: return launch_wgrad<64>(tdy, tx, g, dw, st);
  }
}

// Train-mode (train = 1: mean / invstd from sum, sumsq over `count` rows) or eval-mode (running statistics) BatchNorm
// backward over [M, C] bf16.  When z is given (the ReLU behind the BatchNorm) dz is masked by (z > 0); the masked
// gradient is also written to dzm when dzm != NULL (the shortcut branch's gradient);
// s1 = sum(dz) = dbeta and s2 = sum(dz * xhat) = dgamma are ACCUMULATED (caller zeroes); dy may alias nothing.
static int bn_bwd_impl(const void* dz, void* dzm, const void* z, const void* y, void* dy, const float* gamma, const float* sum,
                       const float* sumsq, const float* running_mean, const float* running_var, float* s1, float* s2, long M,
                       int C, long count, float eps, int train, float hi, void* stream, const char* who) {
  B2_ARG_CHECK(dz && y && dy && gamma && s1 && s2 && M > 0, "%s: null pointer or empty", who);
  B2_ARG_CHECK(dzm == nullptr || z != nullptr, "%s: the masked-gradient output dzm needs the ReLU mask z", who);
  cudaStream_t st = (cudaStream_t)stream;
  const float inv = train ? 1.f / (float)count : 0.f;
  const int groups = C / 8;
  const int lanes = 256 / groups > 0 ? 256 / groups : 1;
  long rpb = (M + (long)b2_num_sms() * 8 - 1) / ((long)b2_num_sms() * 8);
  rpb = (rpb + lanes - 1) / lanes * lanes;
  const unsigned blocks = (unsigned)((M + rpb - 1) / rpb);
  bn_bwd_reduce_kernel<<<blocks, 256, 0, st>>>((const bf16*)dz, (bf16*)dzm, (const bf16*)z, (const bf16*)y, sum, sumsq, running_mean,
                                               running_var, s1, s2, M, C, (int)rpb, inv, eps, train, hi);
  B2_LAUNCH_CHECK("bn_bwd_reduce_kernel");
  // the grid stride (blocks * 256 threads) must be a multiple of the channel-group count for the hoisted coefficients
  unsigned ab = ew_blocks(M * groups);
  if (256 % groups != 0) ab = ab / groups * groups > 0 ? ab / groups * groups : groups;
  bn_bwd_apply_kernel<<<ab, 256, 0, st>>>((const bf16*)dz, (const bf16*)z, (const bf16*)y, (bf16*)dy, gamma, sum, sumsq,
                                          running_mean, running_var, s1, s2, M, C, inv, eps, train, hi);
  B2_LAUNCH_CHECK("bn_bwd_apply_kernel");
  return 0;
}

B2_API int b2_bn_bwd_nhwc_bf16(const void* dz, void* dzm, const void* z, const void* y, void* dy, const float* gamma, const float* sum,
                               const float* sumsq, const float* running_mean, const float* running_var, float* s1, float* s2,
                               long M, int C, long count, float eps, int train, void* stream) {
  return bn_bwd_impl(dz, dzm, z, y, dy, gamma, sum, sumsq, running_mean, running_var, s1, s2, M, C, count, eps, train, INFINITY,
                     stream, "b2_bn_bwd_nhwc_bf16");
}

// the same with the ReLU6 mask 0 < z < 6 (MobileNetV2: nn.ReLU6 behind every BatchNorm but the linear bottleneck's)
B2_API int b2_bn_bwd_relu6_nhwc_bf16(const void* dz, void* dzm, const void* z, const void* y, void* dy, const float* gamma,
                                     const float* sum, const float* sumsq, const float* running_mean, const float* running_var,
                                     float* s1, float* s2, long M, int C, long count, float eps, int train, void* stream) {
  return bn_bwd_impl(dz, dzm, z, y, dy, gamma, sum, sumsq, running_mean, running_var, s1, s2, M, C, count, eps, train, 6.0f, stream,
                     "b2_bn_bwd_relu6_nhwc_bf16");
}

// z [N,H,W,C] (caller zeroes) <- dy [N,P,Q,C] at the even positions: the stride-2 data gradient runs a stride-1 conv over z
B2_API int b2_dilate2_nhwc_bf16(const void* dy, void* z, int N, int P, int Q, int H, int W, int C, void* stream) {
  B2_ARG_CHECK(dy && z && N > 0 && P > 0 && Q > 0 && C % 8 == 0, "b2_dilate2_nhwc_bf16: bad arguments");
  B2_ARG_CHECK(2 * (P - 1) < H && 2 * (Q - 1) < W, "b2_dilate2_nhwc_bf16: dy does not fit the dilated grid");
  dilate2_kernel<<<ew_blocks((long)N * P * Q * (C / 8)), 256, 0, (cudaStream_t)stream>>>((const bf16*)dy, (bf16*)z, N, P, Q,
                                                                                          H, W, C);
  B2_LAUNCH_CHECK("dilate2_kernel");
  return 0;
}

// wk / wt may be NULL (only the other layout is written); see the kernel
B2_API int b2_conv_weight_layouts(const float* w, void* wk, void* wt, int Cout, int Cin, int R, int S, void* stream) {
  B2_ARG_CHECK(w && (wk || wt) && Cout > 0 && Cin > 0 && R > 0 && S > 0, "b2_conv_weight_layouts: bad arguments");
  conv_weight_layouts_kernel<<<ew_blocks((long)Cout * Cin * R * S), 256, 0, (cudaStream_t)stream>>>(w, (bf16*)wk, (bf16*)wt, Cout,
                                                                                                   Cin, R, S);
  B2_LAUNCH_CHECK("conv_weight_layouts_kernel");
  return 0;
}

B2_API int b2_avgpool_bwd_nhwc(const float* dfeat, void* dz, long N, int HW, int C, void* stream) {
  B2_ARG_CHECK(dfeat && dz && N > 0 && HW > 0 && C > 0, "b2_avgpool_bwd_nhwc: bad arguments");
  avgpool_bwd_kernel<<<ew_blocks(N * HW * C), 256, 0, (cudaStream_t)stream>>>(dfeat, (bf16*)dz, N, HW, C);
  B2_LAUNCH_CHECK("avgpool_bwd_kernel");
  return 0;
}

// dbn [N,H,W,C] fp32 (caller zeroes) += routed gradient of maxpool3x3s2p1(relu(raw * scale + shift)); see the kernel
B2_API int b2_maxpool_relu_bwd_nhwc(const void* raw, const float* scale, const float* shift, const void* dpool, float* dbn,
                                    int N, int H, int W, int P, int Q, int C, void* stream) {
  B2_ARG_CHECK(raw && scale && shift && dpool && dbn && N > 0 && H > 0 && W > 0 && P > 0 && Q > 0 && C > 0,
               "b2_maxpool_relu_bwd_nhwc: bad arguments");
  maxpool_relu_bwd_kernel<<<ew_blocks((long)N * P * Q * C), 256, 0, (cudaStream_t)stream>>>(
      (const bf16*)raw, scale, shift, (const bf16*)dpool, dbn, N, H, W, P, Q, C);
  B2_LAUNCH_CHECK("maxpool_relu_bwd_kernel");
  return 0;
}
