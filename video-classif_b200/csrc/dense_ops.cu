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
        const float o = fmaf(a[j], sc[j], sh[j]);
        a[j] = relu ? fmaxf(o, 0.f) : o;
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
