// Element kernels of the DenseNet frame encoder (torchvision densenet121/169/201 under lrcn/lrcn.py:196-209,
// lrcn/rgb_lrcn.py:180-193: `getattr(models, CONF_CNN_BACKBONE)`, classifier -> Identity).
//
// A dense block keeps ONE channel-concatenated NHWC buffer X [pixels, C_final] (row stride C_final): layer k reads the
// first C_k channels and appends `growth` new ones, so torch.cat never copies anything.  The per-channel batch statistics
// of a feature are those of the raw conv output that produced it (taken once, in the producing conv's epilogue, and kept
// in a [2, C_final] table); every BatchNorm that later reads the feature only finalises its own (gamma, beta) against them.
//   * scale_shift_apply_ld_kernel  y[r, :C] = act(x[r, :C] * scale + shift) between row-strided bf16 tensors (the
//                                  pre-activation BN+ReLU of a dense layer / transition; scale = NULL: plain slice copy)
//   * colstats_ld_kernel           per-channel sum / sum of squares of a row-strided bf16 tensor (block input features)
//   * avgpool2x2_kernel            AvgPool2d(2, 2) of a transition into the next block's buffer (row stride ldy)
#include "common.cuh"

namespace {

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
__device__ __forceinline__ uint4 pack8(const float (&o)[8]) {
  uint4 r;
  r.x = pack2(o[0], o[1]); r.y = pack2(o[2], o[3]); r.z = pack2(o[4], o[5]); r.w = pack2(o[6], o[7]);
  return r;
}

__global__ void __launch_bounds__(256)
scale_shift_apply_ld_kernel(const bf16* __restrict__ x, long ldx, bf16* __restrict__ y, long ldy, long rows, int C,
                            const float* __restrict__ scale, const float* __restrict__ shift, int relu) {
  const int groups = C >> 3;
  const long total = rows * groups;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c0 = (int)(i % groups) << 3;
    const long r = i / groups;
    uint4 u = *reinterpret_cast<const uint4*>(x + r * ldx + c0);
    if (scale != nullptr) {
      float a[8];
      unpack8(u, a);
      const float4 s0 = __ldg(reinterpret_cast<const float4*>(scale + c0)), s1 = __ldg(reinterpret_cast<const float4*>(scale + c0 + 4));
      const float4 h0 = __ldg(reinterpret_cast<const float4*>(shift + c0)), h1 = __ldg(reinterpret_cast<const float4*>(shift + c0 + 4));
      const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
      const float sh[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float o = fmaf(a[j], sc[j], sh[j]);
        if (relu >= 1) o = fmaxf(o, 0.f);
        if (relu == 2) o = fminf(o, 6.f);       // ReLU6 (MobileNetV2)
        a[j] = o;
      }
      u = pack8(a);
    }
    *reinterpret_cast<uint4*>(y + r * ldy + c0) = u;
  }
}

// thread = 8 channels, fixed over its rows; block combine through shared memory; one atomic per channel and block
__global__ void __launch_bounds__(256)
colstats_ld_kernel(const bf16* __restrict__ x, long ld, long rows, int C, int rows_per_block, float* __restrict__ sum,
                   float* __restrict__ sumsq) {
  __shared__ float red[2][256 * 8];
  const int groups = C >> 3;
  const int lanes = 256 / groups > 0 ? 256 / groups : 1;
  const int grp = threadIdx.x % groups;
  const int rl = threadIdx.x / groups;
  const bool active = rl < lanes;
  float s1[8], s2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s1[j] = s2[j] = 0.f;
  const long r0 = (long)blockIdx.x * rows_per_block;
  const long r1 = min(rows, r0 + rows_per_block);
  if (active) {
    for (long r = r0 + rl; r < r1; r += lanes) {
      float a[8];
      unpack8(*reinterpret_cast<const uint4*>(x + r * ld + grp * 8), a);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s1[j] += a[j];
        s2[j] = fmaf(a[j], a[j], s2[j]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    red[0][threadIdx.x * 8 + j] = active ? s1[j] : 0.f;
    red[1][threadIdx.x * 8 + j] = active ? s2[j] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += 256) {
    const int gq = i >> 3, j = i & 7;
    float a = 0.f, b = 0.f;
    for (int l = 0; l < lanes; ++l) {
      a += red[0][(l * groups + gq) * 8 + j];
      b += red[1][(l * groups + gq) * 8 + j];
    }
    atomicAdd(sum + i, a);
    atomicAdd(sumsq + i, b);
  }
}

// y[n, p, q, :C] (row stride ldy) = mean of the 2x2 window of x [N, H, W, C]; P = H / 2, Q = W / 2 (floor)
__global__ void __launch_bounds__(256)
avgpool2x2_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, long ldy, int N, int H, int W, int C) {
  const int groups = C >> 3;
  const int P = H >> 1, Q = W >> 1;
  const long total = (long)N * P * Q * groups;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c0 = (int)(i % groups) << 3;
    long px = i / groups;
    const long orow = px;
    const int q = (int)(px % Q);
    px /= Q;
    const int p = (int)(px % P);
    const long n = px / P;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
    for (int dh = 0; dh < 2; ++dh)
#pragma unroll
      for (int dw = 0; dw < 2; ++dw) {
        float a[8];
        unpack8(*reinterpret_cast<const uint4*>(x + (((n * H + 2 * p + dh) * W + 2 * q + dw) * C + c0)), a);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += a[j];
      }
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] *= 0.25f;
    *reinterpret_cast<uint4*>(y + orow * ldy + c0) = pack8(acc);
  }
}

// ---- BatchNorm (+ReLU) backward between row-strided tensors, for the pre-activation BatchNorms of a dense block:
// x = X[:, :C] of the concatenated buffer (stride ldx), dz / z contiguous-ish (their own strides), the result is written or
// ACCUMULATED into the block's gradient buffer dX[:, :C] (stride lddx): every later layer adds its contribution there.
__device__ __forceinline__ void mean_invstd(const float* sum, const float* sumsq, const float* rmean, const float* rvar, int c,
                                            float inv_count, float eps, int train, float& mean, float& invstd) {
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
bn_bwd_reduce_ld_kernel(const bf16* __restrict__ dz, long lddz, const bf16* __restrict__ z, long ldz,
                        const bf16* __restrict__ x, long ldx, const float* __restrict__ sum, const float* __restrict__ sumsq,
                        const float* __restrict__ rmean, const float* __restrict__ rvar, float* __restrict__ out_s1,
                        float* __restrict__ out_s2, long M, int C, int rows_per_block, float inv_count, float eps, int train) {
  __shared__ float red[2][256 * 8];
  const int groups = C >> 3;
  const int lanes = 256 / groups > 0 ? 256 / groups : 1;
  const int grp = threadIdx.x % groups;
  const int rl = threadIdx.x / groups;
  const bool active = rl < lanes;
  float mean[8], invstd[8], s1[8], s2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    s1[j] = s2[j] = 0.f;
    mean[j] = invstd[j] = 0.f;
    if (active) mean_invstd(sum, sumsq, rmean, rvar, grp * 8 + j, inv_count, eps, train, mean[j], invstd[j]);
  }
  const long r0 = (long)blockIdx.x * rows_per_block;
  const long r1 = min(M, r0 + rows_per_block);
  if (active) {
    for (long r = r0 + rl; r < r1; r += lanes) {
      float gf[8], xf[8];
      unpack8(*reinterpret_cast<const uint4*>(dz + r * lddz + grp * 8), gf);
      unpack8(*reinterpret_cast<const uint4*>(x + r * ldx + grp * 8), xf);
      if (z != nullptr) {
        float zf[8];
        unpack8(*reinterpret_cast<const uint4*>(z + r * ldz + grp * 8), zf);
#pragma unroll
        for (int j = 0; j < 8; ++j) gf[j] = zf[j] > 0.f ? gf[j] : 0.f;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s1[j] += gf[j];
        s2[j] = fmaf(gf[j], (xf[j] - mean[j]) * invstd[j], s2[j]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    red[0][threadIdx.x * 8 + j] = active ? s1[j] : 0.f;
    red[1][threadIdx.x * 8 + j] = active ? s2[j] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += 256) {
    const int gq = i >> 3, j = i & 7;
    float a = 0.f, b = 0.f;
    for (int l = 0; l < lanes; ++l) {
      a += red[0][(l * groups + gq) * 8 + j];
      b += red[1][(l * groups + gq) * 8 + j];
    }
    atomicAdd(out_s1 + i, a);
    atomicAdd(out_s2 + i, b);
  }
}

// dx = A[c] dz + B[c] x + K[c] (train; see bn_bwd_apply_kernel in conv_bwd.cu) or A[c] dz (eval).  The launch makes the grid
// stride a multiple of the channel-group count, so a thread's 8 channels -- and its three coefficient vectors -- are fixed.
__global__ void __launch_bounds__(256)
bn_bwd_apply_ld_kernel(const bf16* __restrict__ dz, long lddz, const bf16* __restrict__ z, long ldz,
                       const bf16* __restrict__ x, long ldx, bf16* __restrict__ dx, long lddx, int accumulate,
                       const float* __restrict__ gamma, const float* __restrict__ sum, const float* __restrict__ sumsq,
                       const float* __restrict__ rmean, const float* __restrict__ rvar, const float* __restrict__ s1,
                       const float* __restrict__ s2, long M, int C, float inv_count, float eps, int train) {
  const int groups = C >> 3;
  const long total = M * groups;
  const long i0 = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int grp = (int)(i0 % groups);
  float ca[8], cb[8], ck[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = grp * 8 + j;
    float mean, invstd;
    mean_invstd(sum, sumsq, rmean, rvar, c, inv_count, eps, train, mean, invstd);
    ca[j] = gamma[c] * invstd;
    cb[j] = train ? -ca[j] * invstd * s2[c] * inv_count : 0.f;
    ck[j] = train ? -ca[j] * s1[c] * inv_count - cb[j] * mean : 0.f;
  }
  for (long i = i0; i < total; i += (long)gridDim.x * blockDim.x) {
    const long r = i / groups;
    float gf[8], xf[8], o[8];
    unpack8(*reinterpret_cast<const uint4*>(dz + r * lddz + grp * 8), gf);
    unpack8(*reinterpret_cast<const uint4*>(x + r * ldx + grp * 8), xf);
    if (z != nullptr) {
      float zf[8];
      unpack8(*reinterpret_cast<const uint4*>(z + r * ldz + grp * 8), zf);
#pragma unroll
      for (int j = 0; j < 8; ++j) gf[j] = zf[j] > 0.f ? gf[j] : 0.f;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = fmaf(ca[j], gf[j], fmaf(cb[j], xf[j], ck[j]));
    bf16* dst = dx + r * lddx + grp * 8;
    if (accumulate) {
      float prev[8];
      unpack8(*reinterpret_cast<const uint4*>(dst), prev);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] += prev[j];
    }
    *reinterpret_cast<uint4*>(dst) = pack8(o);
  }
}

// backward of AvgPool2d(2, 2): dx[n, h, w, :] = 0.25 * dy[n, h/2, w/2, :] (rows / columns past 2P, 2Q get zero); dy row-strided
__global__ void __launch_bounds__(256)
avgpool2x2_bwd_kernel(const bf16* __restrict__ dy, long lddy, bf16* __restrict__ dx, int N, int H, int W, int C) {
  const int groups = C >> 3;
  const int P = H >> 1, Q = W >> 1;
  const long total = (long)N * H * W * groups;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c0 = (int)(i % groups) << 3;
    long px = i / groups;
    const long orow = px;
    const int w = (int)(px % W);
    px /= W;
    const int h = (int)(px % H);
    const long n = px / H;
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = 0.f;
    if ((h >> 1) < P && (w >> 1) < Q) {
      unpack8(*reinterpret_cast<const uint4*>(dy + ((n * P + (h >> 1)) * Q + (w >> 1)) * lddy + c0), o);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] *= 0.25f;
    }
    *reinterpret_cast<uint4*>(dx + orow * C + c0) = pack8(o);
  }
}

unsigned ew_blocks(long n) {
  long b = (n + 255) / 256;
  const long cap = (long)b2_num_sms() * 16;
  return (unsigned)(b < cap ? (b > 0 ? b : 1) : cap);
}

}  // namespace

B2_API int b2_scale_shift_apply_ld_bf16(const void* x, long ldx, void* y, long ldy, long rows, int C, const float* scale,
                                        const float* shift, int relu, void* stream) {
  B2_ARG_CHECK(x && y && rows > 0 && C > 0, "b2_scale_shift_apply_ld_bf16: null pointer or empty");
  B2_ARG_CHECK(C % 8 == 0 && ldx % 8 == 0 && ldy % 8 == 0 && ldx >= C && ldy >= C &&
                   ((uintptr_t)x & 15) == 0 && ((uintptr_t)y & 15) == 0,
               "b2_scale_shift_apply_ld_bf16: C and the row strides must be multiples of 8 (16-byte vectors)");
  B2_ARG_CHECK((scale == nullptr) == (shift == nullptr), "b2_scale_shift_apply_ld_bf16: scale and shift go together");
  scale_shift_apply_ld_kernel<<<ew_blocks(rows * (C / 8)), 256, 0, (cudaStream_t)stream>>>(
      (const bf16*)x, ldx, (bf16*)y, ldy, rows, C, scale, shift, relu);
  B2_LAUNCH_CHECK("scale_shift_apply_ld_kernel");
  return 0;
}

// sum / sumsq [C] are ACCUMULATED (caller zeroes)
B2_API int b2_colstats_ld_bf16(const void* x, long ld, long rows, int C, float* sum, float* sumsq, void* stream) {
  B2_ARG_CHECK(x && sum && sumsq && rows > 0, "b2_colstats_ld_bf16: null pointer or empty");
  B2_ARG_CHECK(C % 8 == 0 && C >= 8 && C <= 2048 && ld % 8 == 0 && ld >= C && ((uintptr_t)x & 15) == 0,
               "b2_colstats_ld_bf16: C must be a multiple of 8 in [8, 2048], ld a multiple of 8");
  const int groups = C / 8;
  const int lanes = 256 / groups > 0 ? 256 / groups : 1;
  long rpb = (rows + (long)b2_num_sms() * 8 - 1) / ((long)b2_num_sms() * 8);
  rpb = (rpb + lanes - 1) / lanes * lanes;
  colstats_ld_kernel<<<(unsigned)((rows + rpb - 1) / rpb), 256, 0, (cudaStream_t)stream>>>((const bf16*)x, ld, rows, C,
                                                                                          (int)rpb, sum, sumsq);
  B2_LAUNCH_CHECK("colstats_ld_kernel");
  return 0;
}

B2_API int b2_avgpool2x2_nhwc_bf16(const void* x, void* y, long ldy, int N, int H, int W, int C, void* stream) {
  B2_ARG_CHECK(x && y && N > 0 && H >= 2 && W >= 2, "b2_avgpool2x2_nhwc_bf16: null pointer or empty");
  B2_ARG_CHECK(C % 8 == 0 && ldy % 8 == 0 && ldy >= C, "b2_avgpool2x2_nhwc_bf16: C and ldy must be multiples of 8");
  avgpool2x2_kernel<<<ew_blocks((long)N * (H / 2) * (W / 2) * (C / 8)), 256, 0, (cudaStream_t)stream>>>(
      (const bf16*)x, (bf16*)y, ldy, N, H, W, C);
  B2_LAUNCH_CHECK("avgpool2x2_kernel");
  return 0;
}

// BatchNorm (+ReLU mask z > 0 when z != NULL) backward between row-strided tensors; s1 (= dbeta), s2 (= dgamma) ACCUMULATED
// (caller zeroes); dx[:, :C] is overwritten (accumulate = 0) or added to (accumulate = 1).
B2_API int b2_bn_bwd_ld_bf16(const void* dz, long lddz, const void* z, long ldz, const void* x, long ldx, void* dx, long lddx,
                             int accumulate, const float* gamma, const float* sum, const float* sumsq,
                             const float* running_mean, const float* running_var, float* s1, float* s2, long M, int C,
                             long count, float eps, int train, void* stream) {
  const char* who = "b2_bn_bwd_ld_bf16";
  B2_ARG_CHECK(dz && x && dx && gamma && s1 && s2 && M > 0, "%s: null pointer or empty", who);
  B2_ARG_CHECK(C % 8 == 0 && C >= 8 && C <= 2048, "%s: C must be a multiple of 8 in [8, 2048] (got %d)", who, C);
  B2_ARG_CHECK(lddz % 8 == 0 && ldx % 8 == 0 && lddx % 8 == 0 && (z == nullptr || ldz % 8 == 0), "%s: row strides must be multiples of 8", who);
  B2_ARG_CHECK(train ? (sum && sumsq && count > 0) : (running_mean && running_var), "%s: statistics missing", who);
  cudaStream_t st = (cudaStream_t)stream;
  const float inv = train ? 1.f / (float)count : 0.f;
  const int groups = C / 8;
  const int lanes = 256 / groups > 0 ? 256 / groups : 1;
  long rpb = (M + (long)b2_num_sms() * 8 - 1) / ((long)b2_num_sms() * 8);
  rpb = (rpb + lanes - 1) / lanes * lanes;
  bn_bwd_reduce_ld_kernel<<<(unsigned)((M + rpb - 1) / rpb), 256, 0, st>>>((const bf16*)dz, lddz, (const bf16*)z, ldz,
                                                                          (const bf16*)x, ldx, sum, sumsq, running_mean,
                                                                          running_var, s1, s2, M, C, (int)rpb, inv, eps, train);
  B2_LAUNCH_CHECK("bn_bwd_reduce_ld_kernel");
  unsigned ab = ew_blocks(M * groups);       // grid stride (ab * 256) must be a multiple of the channel-group count
  if (256 % groups != 0) ab = ab / groups * groups > 0 ? ab / groups * groups : (unsigned)groups;
  bn_bwd_apply_ld_kernel<<<ab, 256, 0, st>>>((const bf16*)dz, lddz, (const bf16*)z, ldz, (const bf16*)x,
                                                               ldx, (bf16*)dx, lddx, accumulate, gamma, sum, sumsq,
                                                               running_mean, running_var, s1, s2, M, C, inv, eps, train);
  B2_LAUNCH_CHECK("bn_bwd_apply_ld_kernel");
  return 0;
}

B2_API int b2_avgpool2x2_bwd_nhwc_bf16(const void* dy, long lddy, void* dx, int N, int H, int W, int C, void* stream) {
  B2_ARG_CHECK(dy && dx && N > 0 && H >= 2 && W >= 2, "b2_avgpool2x2_bwd_nhwc_bf16: null pointer or empty");
  B2_ARG_CHECK(C % 8 == 0 && lddy % 8 == 0 && lddy >= C, "b2_avgpool2x2_bwd_nhwc_bf16: C and lddy must be multiples of 8");
  avgpool2x2_bwd_kernel<<<ew_blocks((long)N * H * W * (C / 8)), 256, 0, (cudaStream_t)stream>>>((const bf16*)dy, lddy,
                                                                                               (bf16*)dx, N, H, W, C);
  B2_LAUNCH_CHECK("avgpool2x2_bwd_kernel");
  return 0;
}
