// NHWC bf16 glue kernels of the frame-CNN (ResNet-class backbone) path -- all HBM-bound,
// 128-bit (8 x bf16) vector accesses, grids sized in multiples of the SM count.
//
//   stem_im2col     : fp32/bf16 NCHW frames -> [M, Kp] bf16 patch matrix for the 7x7/2 stem GEMM
//   bn_apply        : train/eval BatchNorm (statistics come from the conv epilogue's column sums)
//                     + optional residual (identity or a second raw tensor with its own BN) + ReLU
//   bn_relu_maxpool : stem BN + ReLU + 3x3/2 max-pool in one pass
//   avgpool         : global average pool -> [N, C] fp32 (+ bf16 copy for the adapt GEMM)
//
// BatchNorm semantics follow torch.nn.BatchNorm2d as the reference runs it (train mode even for
// frozen backbones: medsos_lrcn/src/train_eval.py:12 + models.py:144-145): biased variance for
// normalisation, running_var updated with the unbiased one, momentum 0.1.
#include "common.cuh"

namespace {

struct BnSrc {
  const float* sum;     // [C] column sums from the producing conv (train) or null (eval)
  const float* sumsq;   // [C]
  const float* gamma;   // [C]
  const float* beta;    // [C]
  float* running_mean;  // [C] updated in train mode, read in eval mode
  float* running_var;   // [C]
};

// scale/shift for channel c ; block 0 also performs the running-stat update
__device__ __forceinline__ float2 bn_scale_shift(const BnSrc& b, int c, float inv_count, float unbias, float eps,
                                                 float momentum, bool train, bool update) {
  float mean, var;
  if (train) {
    mean = b.sum[c] * inv_count;
    var = fmaxf(b.sumsq[c] * inv_count - mean * mean, 0.f);
    if (update && b.running_mean != nullptr) {
      b.running_mean[c] = (1.f - momentum) * b.running_mean[c] + momentum * mean;
      b.running_var[c] = (1.f - momentum) * b.running_var[c] + momentum * var * unbias;
    }
  } else {
    mean = b.running_mean[c];
    var = b.running_var[c];
  }
  const float rstd = rsqrtf(var + eps);
  const float sc = b.gamma[c] * rstd;
  return make_float2(sc, b.beta[c] - mean * sc);
}

__device__ __forceinline__ void load8(const bf16* p, float (&v)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  unpack_bf16x2(u.x, v[0], v[1]);
  unpack_bf16x2(u.y, v[2], v[3]);
  unpack_bf16x2(u.z, v[4], v[5]);
  unpack_bf16x2(u.w, v[6], v[7]);
}
__device__ __forceinline__ void store8(bf16* p, const float (&v)[8]) {
  uint4 u;
  u.x = pack_bf16x2(v[0], v[1]);
  u.y = pack_bf16x2(v[2], v[3]);
  u.z = pack_bf16x2(v[4], v[5]);
  u.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(p) = u;
}

// res_mode 0: none, 1: residual already normalised (identity shortcut), 2: residual raw + own BN
// Fast path (C/8 divides the block size, true for every ResNet width): a thread's 8 channels never
// change while it grid-strides, so their scale/shift live in registers -- the loop is pure
// load / fma / store with U independent 16-byte loads in flight per stream.
template <int RES_MODE>
__global__ void __launch_bounds__(256)
bn_apply_reg_kernel(const bf16* x, bf16* y, long rows, int C, BnSrc bn, const bf16* __restrict__ res, BnSrc rbn,
                    float inv_count, float unbias, float eps, float momentum, int train, int relu) {
  const int vec_per_row = C >> 3;
  const int c0 = (threadIdx.x % vec_per_row) << 3;
  const bool updater = blockIdx.x == 0 && threadIdx.x < vec_per_row;
  float sc[8], sh[8], sc2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float2 s = bn_scale_shift(bn, c0 + j, inv_count, unbias, eps, momentum, train, updater);
    sc[j] = s.x;
    sh[j] = s.y;
    if (RES_MODE == 2) {
      const float2 s2 = bn_scale_shift(rbn, c0 + j, inv_count, unbias, eps, momentum, train, updater);
      sc2[j] = s2.x;
      sh[j] += s2.y;
    }
  }
  const long total = rows * vec_per_row;
  const long stride = (long)gridDim.x * blockDim.x;     // multiple of vec_per_row
  constexpr int U = 4;
  for (long v0 = (long)blockIdx.x * blockDim.x + threadIdx.x; v0 < total; v0 += stride * U) {
    uint4 xa[U], ra[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long v = v0 + u * stride;
      if (v < total) {
        xa[u] = *reinterpret_cast<const uint4*>(x + v * 8);
        if (RES_MODE) ra[u] = __ldg(reinterpret_cast<const uint4*>(res + v * 8));
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long v = v0 + u * stride;
      if (v >= total) break;
      float a[8], r[8];
      unpack_bf16x2(xa[u].x, a[0], a[1]);
      unpack_bf16x2(xa[u].y, a[2], a[3]);
      unpack_bf16x2(xa[u].z, a[4], a[5]);
      unpack_bf16x2(xa[u].w, a[6], a[7]);
      if (RES_MODE) {
        unpack_bf16x2(ra[u].x, r[0], r[1]);
        unpack_bf16x2(ra[u].y, r[2], r[3]);
        unpack_bf16x2(ra[u].z, r[4], r[5]);
        unpack_bf16x2(ra[u].w, r[6], r[7]);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float o = fmaf(a[j], sc[j], sh[j]);
        if (RES_MODE == 1) o += r[j];
        if (RES_MODE == 2) o = fmaf(r[j], sc2[j], o);
        a[j] = relu ? fmaxf(o, 0.f) : o;
      }
      store8(y + v * 8, a);
    }
  }
}

// y = act( x * scale[c] + shift[c] [+ res | + res * rscale[c] + rshift[c]] ) with PRECOMPUTED per-channel
// scale/shift (written by the BatchNorm finalisation tail of the producing conv kernel, gemm_tc.cu).
// Any C multiple of 8: the channel group of a vector is recomputed per element (one integer modulo).
template <int RES_MODE>
__global__ void __launch_bounds__(256)
scale_shift_apply_kernel(const bf16* x, bf16* y, long rows, int C, const float* __restrict__ scale,
                         const float* __restrict__ shift, const bf16* __restrict__ res,
                         const float* __restrict__ rscale, const float* __restrict__ rshift, int relu) {
  const int vec_per_row = C >> 3;
  const long total = rows * vec_per_row;
  for (long v = (long)blockIdx.x * blockDim.x + threadIdx.x; v < total; v += (long)gridDim.x * blockDim.x) {
    const int c0 = (int)(v % vec_per_row) << 3;
    float a[8], r[8];
    load8(x + v * 8, a);
    if (RES_MODE) load8(res + v * 8, r);
    const float4 s0 = __ldg(reinterpret_cast<const float4*>(scale + c0)), s1 = __ldg(reinterpret_cast<const float4*>(scale + c0 + 4));
    const float4 h0 = __ldg(reinterpret_cast<const float4*>(shift + c0)), h1 = __ldg(reinterpret_cast<const float4*>(shift + c0 + 4));
    const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
    const float sh[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
    float rs[8], rh[8];
    if (RES_MODE == 2) {
      const float4 t0 = __ldg(reinterpret_cast<const float4*>(rscale + c0)), t1 = __ldg(reinterpret_cast<const float4*>(rscale + c0 + 4));
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(rshift + c0)), g1 = __ldg(reinterpret_cast<const float4*>(rshift + c0 + 4));
      rs[0] = t0.x; rs[1] = t0.y; rs[2] = t0.z; rs[3] = t0.w; rs[4] = t1.x; rs[5] = t1.y; rs[6] = t1.z; rs[7] = t1.w;
      rh[0] = g0.x; rh[1] = g0.y; rh[2] = g0.z; rh[3] = g0.w; rh[4] = g1.x; rh[5] = g1.y; rh[6] = g1.z; rh[7] = g1.w;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float o = fmaf(a[j], sc[j], sh[j]);
      if (RES_MODE == 1) o += r[j];
      if (RES_MODE == 2) o += fmaf(r[j], rs[j], rh[j]);
      a[j] = relu ? fmaxf(o, 0.f) : o;
    }
    store8(y + v * 8, a);
  }
}

// generic fallback (any C multiple of 8): scale/shift staged in shared memory
__global__ void __launch_bounds__(256)
bn_apply_kernel(const bf16* x, bf16* y, long rows, int C, BnSrc bn, int res_mode,
                const bf16* __restrict__ res, BnSrc rbn, float inv_count, float unbias, float eps, float momentum,
                int train, int relu) {
  extern __shared__ float2 ss[];  // [C] (+ [C] for the residual BN)
  float2* ss2 = ss + C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    ss[c] = bn_scale_shift(bn, c, inv_count, unbias, eps, momentum, train, blockIdx.x == 0);
    if (res_mode == 2) ss2[c] = bn_scale_shift(rbn, c, inv_count, unbias, eps, momentum, train, blockIdx.x == 0);
  }
  __syncthreads();
  const int vec_per_row = C >> 3;
  const long total = rows * vec_per_row;
  for (long v = (long)blockIdx.x * blockDim.x + threadIdx.x; v < total; v += (long)gridDim.x * blockDim.x) {
    const int c0 = (int)(v % vec_per_row) << 3;
    float a[8], r[8];
    load8(x + v * 8, a);
    if (res_mode) load8(res + v * 8, r);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float2 s = ss[c0 + j];
      float o = fmaf(a[j], s.x, s.y);
      if (res_mode == 1) o += r[j];
      else if (res_mode == 2) {
        const float2 s2 = ss2[c0 + j];
        o += fmaf(r[j], s2.x, s2.y);
      }
      a[j] = relu ? fmaxf(o, 0.f) : o;
    }
    store8(y + v * 8, a);
  }
}

__global__ void __launch_bounds__(256)
bn_relu_maxpool_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, int N, int H, int W, int C, int P, int Q,
                       BnSrc bn, float inv_count, float unbias, float eps, float momentum, int train) {
  extern __shared__ float2 ss[];
  for (int c = threadIdx.x; c < C; c += blockDim.x)
    ss[c] = bn_scale_shift(bn, c, inv_count, unbias, eps, momentum, train, blockIdx.x == 0);
  __syncthreads();
  const int vec = C >> 3;
  const long total = (long)N * P * Q * vec;
  for (long v = (long)blockIdx.x * blockDim.x + threadIdx.x; v < total; v += (long)gridDim.x * blockDim.x) {
    const int cg = (int)(v % vec);
    long t = v / vec;
    const int q = (int)(t % Q);
    t /= Q;
    const int p = (int)(t % P);
    const int n = (int)(t / P);
    float m[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) m[j] = 0.f;  // post-ReLU values are >= 0 and the window is never empty
    uint4 win[9];
    bool ok[9];
#pragma unroll
    for (int t9 = 0; t9 < 9; ++t9) {        // issue all 9 window loads before using any
      const int iy = 2 * p - 1 + t9 / 3, ix = 2 * q - 1 + t9 % 3;
      ok[t9] = iy >= 0 && iy < H && ix >= 0 && ix < W;
      if (ok[t9]) win[t9] = __ldg(reinterpret_cast<const uint4*>(x + (((long)n * H + iy) * W + ix) * C + cg * 8));
    }
#pragma unroll
    for (int t9 = 0; t9 < 9; ++t9) {
      if (!ok[t9]) continue;
      float a[8];
      unpack_bf16x2(win[t9].x, a[0], a[1]);
      unpack_bf16x2(win[t9].y, a[2], a[3]);
      unpack_bf16x2(win[t9].z, a[4], a[5]);
      unpack_bf16x2(win[t9].w, a[6], a[7]);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float2 sc = ss[cg * 8 + j];
        m[j] = fmaxf(m[j], fmaf(a[j], sc.x, sc.y));
      }
    }
    store8(y + v * 8, m);
  }
}


// Fast path of the stem BN + ReLU + 3x3/2 max-pool (C/8 divides 256, true for the 64-channel stem).
// BN is monotone per channel and ReLU is monotone, so max_t relu(x_t*sc + sh) = relu(ext*sc + sh) with
// ext = max_t x_t (sc >= 0) or min_t x_t (sc < 0): the window is reduced on the RAW bf16 values with packed
// bf16x2 max / min (exact, 8 instructions per tap instead of 32) and the affine runs once per output -- the
// kernel was ALU-issue bound (ncu: IPC 2.4, 128 M warp instructions) with one fma + max per tap and channel.
// Out-of-image taps contribute -inf / +inf; the window always holds at least one real pixel.
__global__ void __launch_bounds__(256)
bn_relu_maxpool_reg_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, int N, int H, int W, int C, int P, int Q,
                           BnSrc bn, float inv_count, float unbias, float eps, float momentum, int train) {
  const int vec = C >> 3;
  const int cg = threadIdx.x % vec;
  const int pix_per_block = 256 / vec;
  const bool updater = blockIdx.x == 0 && threadIdx.x < vec;
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float2 s = bn_scale_shift(bn, cg * 8 + j, inv_count, unbias, eps, momentum, train, updater);
    sc[j] = s.x;
    sh[j] = s.y;
  }
  // sign trick: max_t relu(x_t sc + sh) = relu(|sc| max_t(sign(sc) x_t) + sh), and flipping a bf16 sign is one XOR on the packed
  // pair -> ONE packed max per word and tap (the min / max pair + selects of the previous version made the kernel ALU bound at
  // 0.48 of the copy bandwidth).  Padding taps are replaced by the clamped (always in-window) pixel: a duplicate never changes a max.
  uint32_t flip[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    flip[k] = (sc[2 * k] < 0.f ? 0x00008000u : 0u) | (sc[2 * k + 1] < 0.f ? 0x80000000u : 0u);
    sc[2 * k] = fabsf(sc[2 * k]);
    sc[2 * k + 1] = fabsf(sc[2 * k + 1]);
  }
  const unsigned npix = (unsigned)((long)N * P * Q);          // host checks < 2^31
  for (unsigned pix = blockIdx.x * pix_per_block + threadIdx.x / vec; pix < npix; pix += gridDim.x * pix_per_block) {
    const unsigned t = pix / (unsigned)Q;
    const int q = (int)(pix - t * (unsigned)Q);
    const unsigned n = t / (unsigned)P;
    const int p = (int)(t - n * (unsigned)P);
    const bf16* base = x + ((long)n * H * W) * C + cg * 8;
    uint4 win[9];
#pragma unroll
    for (int t9 = 0; t9 < 9; ++t9) {           // all 9 window loads in flight before any use
      const int iy = min(max(2 * p - 1 + t9 / 3, 0), H - 1), ix = min(max(2 * q - 1 + t9 % 3, 0), W - 1);
      win[t9] = __ldg(reinterpret_cast<const uint4*>(base + ((long)iy * W + ix) * C));
    }
    uint32_t mx[4] = {win[0].x ^ flip[0], win[0].y ^ flip[1], win[0].z ^ flip[2], win[0].w ^ flip[3]};
#pragma unroll
    for (int t9 = 1; t9 < 9; ++t9) {
      const uint32_t w4[4] = {win[t9].x ^ flip[0], win[t9].y ^ flip[1], win[t9].z ^ flip[2], win[t9].w ^ flip[3]};
#pragma unroll
      for (int k = 0; k < 4; ++k) asm("max.bf16x2 %0, %1, %2;" : "=r"(mx[k]) : "r"(mx[k]), "r"(w4[k]));
    }
    float m[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float v0, v1;
      unpack_bf16x2(mx[k], v0, v1);
      m[2 * k] = fmaxf(fmaf(v0, sc[2 * k], sh[2 * k]), 0.f);
      m[2 * k + 1] = fmaxf(fmaf(v1, sc[2 * k + 1], sh[2 * k + 1]), 0.f);
    }
    store8(y + (long)pix * C + cg * 8, m);
  }
}

// one thread per (n, 8-channel group); HW <= a few hundred
__global__ void __launch_bounds__(256)
avgpool_kernel(const bf16* __restrict__ x, float* __restrict__ out_f32, bf16* __restrict__ out_bf16, int N, int HW,
               int C) {
  const int vec = C >> 3;
  const long total = (long)N * vec;
  for (long v = (long)blockIdx.x * blockDim.x + threadIdx.x; v < total; v += (long)gridDim.x * blockDim.x) {
    const int cg = (int)(v % vec);
    const long n = v / vec;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    const bf16* p = x + n * HW * C + cg * 8;
    for (int i = 0; i < HW; ++i) {
      float a[8];
      load8(p + (long)i * C, a);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += a[j];
    }
    const float inv = 1.f / (float)HW;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] *= inv;
    if (out_f32) {
      float* o = out_f32 + n * C + cg * 8;
      *reinterpret_cast<float4*>(o) = make_float4(acc[0], acc[1], acc[2], acc[3]);
      *reinterpret_cast<float4*>(o + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
    }
    if (out_bf16) store8(out_bf16 + n * C + cg * 8, acc);
  }
}

// Patch matrix for the 7x7 stride-2 pad-3 stem: row m = (n,p,q), column k = (c*7 + r)*8 + s with the
// filter row padded from 7 to 8 taps (the weight matrix carries a zero in the s = 7 slot), so each
// 16-byte output vector is 8 CONSECUTIVE input pixels of one input row: contiguous reads, no
// per-element div/mod.  Kp = 3*7*8 = 168.
template <typename InT>
__global__ void __launch_bounds__(256)
stem_im2col_kernel(const InT* __restrict__ x, bf16* __restrict__ A, int N, int H, int W, int P, int Q, int Kp) {
  const int kvec = Kp >> 3;   // 21
  const long total = (long)N * P * Q * kvec;
  const long plane = (long)H * W;
  for (long v = (long)blockIdx.x * blockDim.x + threadIdx.x; v < total; v += (long)gridDim.x * blockDim.x) {
    const int kv = (int)(v % kvec);
    long m = v / kvec;
    const int q = (int)(m % Q);
    long t = m / Q;
    const int p = (int)(t % P);
    const int n = (int)(t / P);
    const int c = kv / 7, r = kv - c * 7;
    const int iy = 2 * p - 3 + r;
    const int ix0 = 2 * q - 3;
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = 0.f;
    if (c < 3 && iy >= 0 && iy < H) {
      const InT* row = x + ((long)n * 3 + c) * plane + (long)iy * W;
#pragma unroll
      for (int j = 0; j < 7; ++j) {
        const int ix = ix0 + j;
        if (ix >= 0 && ix < W) o[j] = (float)row[ix];
      }
    }
    store8(A + v * 8, o);
  }
}

int grid_for(long work_items, int threads) {
  long blocks = (work_items + threads - 1) / threads;
  long cap = (long)b2_num_sms() * 16;
  return (int)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

}  // namespace

B2_API int b2_bn_apply_nhwc(const void* x, void* y, long rows, int C, const float* sum, const float* sumsq,
                            const float* gamma, const float* beta, float* running_mean, float* running_var,
                            int res_mode, const void* res, const float* rsum, const float* rsumsq,
                            const float* rgamma, const float* rbeta, float* rrunning_mean, float* rrunning_var,
                            long count, float eps, float momentum, int train, int relu, void* stream) {
  B2_ARG_CHECK(x && y && gamma && beta && rows > 0, "b2_bn_apply_nhwc: null pointer or empty");
  B2_ARG_CHECK(C % 8 == 0 && C <= 4096, "b2_bn_apply_nhwc: C must be a multiple of 8 and <= 4096 (got %d)", C);
  B2_ARG_CHECK(train ? (sum && sumsq) : (running_mean && running_var), "b2_bn_apply_nhwc: missing statistics");
  B2_ARG_CHECK(res_mode == 0 || res, "b2_bn_apply_nhwc: residual pointer missing");
  B2_ARG_CHECK(res_mode != 2 || (rgamma && rbeta && (train ? (rsum && rsumsq) : (rrunning_mean && rrunning_var))),
               "b2_bn_apply_nhwc: residual BN parameters missing");
  BnSrc bn = {sum, sumsq, gamma, beta, running_mean, running_var};
  BnSrc rbn = {rsum, rsumsq, rgamma, rbeta, rrunning_mean, rrunning_var};
  const size_t smem = (size_t)C * sizeof(float2) * (res_mode == 2 ? 2 : 1);
  if (smem > 48 * 1024)
    B2_CUDA_CHECK(cudaFuncSetAttribute(bn_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  const float inv = 1.f / (float)count;
  const float unbias = count > 1 ? (float)((double)count / (double)(count - 1)) : 1.f;
  const int vec_per_row = C / 8;
  cudaStream_t st = (cudaStream_t)stream;
  if (vec_per_row <= 256 && 256 % vec_per_row == 0) {
    const int grid = grid_for((rows * vec_per_row + 3) / 4, 256);
    if (res_mode == 0)
      bn_apply_reg_kernel<0><<<grid, 256, 0, st>>>((const bf16*)x, (bf16*)y, rows, C, bn, (const bf16*)res, rbn, inv,
                                                   unbias, eps, momentum, train, relu);
    else if (res_mode == 1)
      bn_apply_reg_kernel<1><<<grid, 256, 0, st>>>((const bf16*)x, (bf16*)y, rows, C, bn, (const bf16*)res, rbn, inv,
                                                   unbias, eps, momentum, train, relu);
    else
      bn_apply_reg_kernel<2><<<grid, 256, 0, st>>>((const bf16*)x, (bf16*)y, rows, C, bn, (const bf16*)res, rbn, inv,
                                                   unbias, eps, momentum, train, relu);
  } else {
    bn_apply_kernel<<<grid_for(rows * vec_per_row, 256), 256, smem, st>>>(
        (const bf16*)x, (bf16*)y, rows, C, bn, res_mode, (const bf16*)res, rbn, inv, unbias, eps, momentum, train, relu);
  }
  B2_LAUNCH_CHECK("bn_apply_kernel");
  return 0;
}

B2_API int b2_scale_shift_apply_nhwc(const void* x, void* y, long rows, int C, const float* scale, const float* shift,
                                     const void* res, const float* rscale, const float* rshift, int relu,
                                     void* stream) {
  B2_ARG_CHECK(x && y && scale && shift && rows > 0, "b2_scale_shift_apply_nhwc: null pointer or empty");
  B2_ARG_CHECK(C % 8 == 0, "b2_scale_shift_apply_nhwc: C must be a multiple of 8 (got %d)", C);
  B2_ARG_CHECK((rscale == nullptr) == (rshift == nullptr) && (rscale == nullptr || res != nullptr),
               "b2_scale_shift_apply_nhwc: shortcut BatchNorm needs rscale, rshift and res");
  const int grid = grid_for(rows * (C / 8), 256);
  cudaStream_t st = (cudaStream_t)stream;
  if (res == nullptr)
    scale_shift_apply_kernel<0><<<grid, 256, 0, st>>>((const bf16*)x, (bf16*)y, rows, C, scale, shift, nullptr, nullptr,
                                                      nullptr, relu);
  else if (rscale == nullptr)
    scale_shift_apply_kernel<1><<<grid, 256, 0, st>>>((const bf16*)x, (bf16*)y, rows, C, scale, shift, (const bf16*)res,
                                                      nullptr, nullptr, relu);
  else
    scale_shift_apply_kernel<2><<<grid, 256, 0, st>>>((const bf16*)x, (bf16*)y, rows, C, scale, shift, (const bf16*)res,
                                                      rscale, rshift, relu);
  B2_LAUNCH_CHECK("scale_shift_apply_kernel");
  return 0;
}

B2_API int b2_bn_relu_maxpool_nhwc(const void* x, void* y, int N, int H, int W, int C, const float* sum,
                                   const float* sumsq, const float* gamma, const float* beta, float* running_mean,
                                   float* running_var, float eps, float momentum, int train, void* stream) {
  B2_ARG_CHECK(x && y && gamma && beta && N > 0 && H > 0 && W > 0, "b2_bn_relu_maxpool_nhwc: null pointer or empty");
  B2_ARG_CHECK(C % 8 == 0 && C <= 4096, "b2_bn_relu_maxpool_nhwc: C must be a multiple of 8 and <= 4096");
  B2_ARG_CHECK(train ? (sum && sumsq) : (running_mean && running_var), "b2_bn_relu_maxpool_nhwc: missing statistics");
  const int P = (H + 2 - 3) / 2 + 1, Q = (W + 2 - 3) / 2 + 1;
  BnSrc bn = {sum, sumsq, gamma, beta, running_mean, running_var};
  const long count = (long)N * H * W;
  const float inv = 1.f / (float)count;
  const float unbias = count > 1 ? (float)((double)count / (double)(count - 1)) : 1.f;
  if (256 % (C / 8) == 0 && (long)N * P * Q < (1L << 31)) {
    bn_relu_maxpool_reg_kernel<<<grid_for((long)N * P * Q * (C / 8), 256), 256, 0, (cudaStream_t)stream>>>(
        (const bf16*)x, (bf16*)y, N, H, W, C, P, Q, bn, inv, unbias, eps, momentum, train);
  } else {
    bn_relu_maxpool_kernel<<<grid_for((long)N * P * Q * (C / 8), 256), 256, (size_t)C * sizeof(float2),
                             (cudaStream_t)stream>>>((const bf16*)x, (bf16*)y, N, H, W, C, P, Q, bn, inv, unbias, eps,
                                                     momentum, train);
  }
  B2_LAUNCH_CHECK("bn_relu_maxpool_kernel");
  return 0;
}

B2_API int b2_avgpool_nhwc(const void* x, float* out_f32, void* out_bf16, int N, int HW, int C, void* stream) {
  B2_ARG_CHECK(x && (out_f32 || out_bf16) && N > 0 && HW > 0, "b2_avgpool_nhwc: null pointer or empty");
  B2_ARG_CHECK(C % 8 == 0, "b2_avgpool_nhwc: C must be a multiple of 8");
  avgpool_kernel<<<grid_for((long)N * (C / 8), 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)x, out_f32,
                                                                                     (bf16*)out_bf16, N, HW, C);
  B2_LAUNCH_CHECK("avgpool_kernel");
  return 0;
}

B2_API int b2_stem_im2col(const void* x, int in_bf16, void* A, int N, int H, int W, int Kp, void* stream) {
  B2_ARG_CHECK(x && A && N > 0 && H > 0 && W > 0, "b2_stem_im2col: null pointer or empty");
  B2_ARG_CHECK(Kp == 168, "b2_stem_im2col: Kp must be 168 (3 channels x 7 rows x 8 padded taps)");
  const int P = (H + 6 - 7) / 2 + 1, Q = (W + 6 - 7) / 2 + 1;
  const long total = (long)N * P * Q * (Kp / 8);
  if (in_bf16)
    stem_im2col_kernel<bf16><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)x, (bf16*)A, N, H, W,
                                                                                     P, Q, Kp);
  else
    stem_im2col_kernel<float><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>((const float*)x, (bf16*)A, N, H,
                                                                                      W, P, Q, Kp);
  B2_LAUNCH_CHECK("stem_im2col_kernel");
  return 0;
}
