// MobileNetV2 frame encoder kernels (torchvision mobilenet_v2 under medsos_lrcn/src/models.py:133-143; it is in the
// reference's own backbone search space, medsos_lrcn/src/automation.py:28).  Every layer of this network is memory bound
// (depthwise 3x3 convs and thin 1x1 convs), so the kernels minimise passes over the activations:
//   * stem3x3s2_kernel   Conv2d(3, 32, 3, stride 2, pad 1) straight from the NCHW fp32 / bf16 frames to NHWC bf16, with
//                        the per-channel batch statistics of its output
//   * dwconv3x3_kernel   depthwise 3x3 (stride 1 / 2, pad 1) over NHWC bf16 with the PREVIOUS BatchNorm + ReLU6 applied to
//                        the input on load (padding stays zero), raw bf16 output + its batch statistics
// The 1x1 convs run on the tcgen05 GEMM (gemm_tc.cu) with the statistics in their epilogue.
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
__device__ __forceinline__ float bf16_round(float v) { return __bfloat162float(__float2bfloat16(v)); }

// block-level combine of per-thread (8 channel) partial sums: threads with the same channel group are summed, then one
// atomic per channel.  red: [2][256 * 8] floats of shared memory.
__device__ __forceinline__ void flush_stats(float (&s1)[8], float (&s2)[8], int grp, int groups, bool active, float* red,
                                            float* __restrict__ sum, float* __restrict__ sumsq) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    red[threadIdx.x * 8 + j] = active ? s1[j] : 0.f;
    red[256 * 8 + threadIdx.x * 8 + j] = active ? s2[j] : 0.f;
  }
  __syncthreads();
  const int C = groups * 8;
  const int base = (int)(((long)blockIdx.x * 256) % groups);       // thread t owns channel group (base + t) % groups
  for (int i = threadIdx.x; i < C; i += 256) {
    const int gq = i >> 3, j = i & 7;
    float a = 0.f, b = 0.f;
    for (int t = (gq - base + groups) % groups; t < 256; t += groups) {
      a += red[t * 8 + j];
      b += red[256 * 8 + t * 8 + j];
    }
    atomicAdd(sum + i, a);
    atomicAdd(sumsq + i, b);
  }
}

// same, for kernels whose thread t owns channel group t % groups
__device__ __forceinline__ void flush_stats_local(float (&s1)[8], float (&s2)[8], int groups, bool active, float* red,
                                                  float* __restrict__ sum, float* __restrict__ sumsq) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    red[threadIdx.x * 8 + j] = active ? s1[j] : 0.f;
    red[256 * 8 + threadIdx.x * 8 + j] = active ? s2[j] : 0.f;
  }
  __syncthreads();
  const int C = groups * 8;
  for (int i = threadIdx.x; i < C; i += 256) {
    const int gq = i >> 3, j = i & 7;
    float a = 0.f, b = 0.f;
    for (int t = gq; t < 256; t += groups) {
      a += red[t * 8 + j];
      b += red[256 * 8 + t * 8 + j];
    }
    atomicAdd(sum + i, a);
    atomicAdd(sumsq + i, b);
  }
}

// thread = (output pixel, 8 of the 32 output channels); grid stride is a multiple of 4 so the channel group is fixed
template <typename T>
__global__ void __launch_bounds__(256)
stem3x3s2_kernel(const T* __restrict__ x, const float* __restrict__ w, bf16* __restrict__ y, float* __restrict__ sum,
                 float* __restrict__ sumsq, int N, int H, int W, int P, int Q) {
  __shared__ float ws[27 * 32];          // [c*9 + r*3 + s][cout], bf16-rounded operands
  __shared__ float red[2 * 256 * 8];
  for (int i = threadIdx.x; i < 27 * 32; i += 256) {
    const int k = i >> 5, co = i & 31;
    ws[i] = bf16_round(w[co * 27 + k]);
  }
  __syncthreads();
  const long total = (long)N * P * Q * 4;
  const long i0 = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int grp = (int)(i0 & 3);
  float s1[8], s2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s1[j] = s2[j] = 0.f;
  for (long i = i0; i < total; i += (long)gridDim.x * blockDim.x) {
    long px = i >> 2;
    const int q = (int)(px % Q);
    px /= Q;
    const int p = (int)(px % P);
    const long n = px / P;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const int h = 2 * p - 1 + r;
        if (h < 0 || h >= H) continue;
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          const int wq = 2 * q - 1 + s;
          if (wq < 0 || wq >= W) continue;
          const float v = bf16_round((float)x[((n * 3 + c) * H + h) * W + wq]);
          const float* wk = ws + (c * 9 + r * 3 + s) * 32 + grp * 8;
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] = fmaf(v, wk[j], acc[j]);
        }
      }
    const uint4 o = pack8(acc);
    *reinterpret_cast<uint4*>(y + (i >> 2) * 32 + grp * 8) = o;
    if (sum != nullptr) {
      float r8[8];
      unpack8(o, r8);                    // statistics of the value as stored
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s1[j] += r8[j];
        s2[j] = fmaf(r8[j], r8[j], s2[j]);
      }
    }
  }
  if (sum != nullptr) flush_stats(s1, s2, grp, 4, true, red, sum, sumsq);
}

// thread = (strip of TW consecutive output pixels of one row, 8 channels); the channel group is fixed per thread (the launch
// makes the grid stride a multiple of the group count), so the 9 x 8 filter taps and the BatchNorm coefficients live in
// registers.  A strip shares its input window: 3 rows x ((TW-1) * STRIDE + 3) pixels are loaded (all loads of a row issued
// before any use), normalised + activated ONCE each and scattered into the <= 3 outputs they feed -- 4.5 (stride 1) / 7.5
// (stride 2) loads and activations per output instead of 9.  act: 0 none, 1 ReLU, 2 ReLU6 on x * scale + shift (scale =
// NULL: x as it is); padding contributes exact zeros.
template <int STRIDE, bool TF>
__global__ void __launch_bounds__(256, 2)
dwconv3x3_kernel(const bf16* __restrict__ x, const float* __restrict__ scale, const float* __restrict__ shift, int act,
                 const float* __restrict__ w, bf16* __restrict__ y, float* __restrict__ sum, float* __restrict__ sumsq, int N,
                 int H, int W, int C, int P, int Q) {
  constexpr int TW = STRIDE == 1 ? 4 : 2;
  constexpr int NIN = (TW - 1) * STRIDE + 3;
  // filter taps [9][C] and BatchNorm coefficients [2][C] in shared memory (72 + 16 registers per thread otherwise: the kernel
  // spilled at two blocks per SM); the same bytes are the statistics scratch after the main loop.  C <= 1024.
  __shared__ __align__(16) float sm[11 * 1024];
  float* wsm = sm;
  float* ssm = sm + 9 * C;
  for (int i = threadIdx.x; i < 9 * C; i += 256) {
    const int t = i / C, c = i - t * C;
    wsm[i] = bf16_round(w[c * 9 + t]);
  }
  constexpr bool tf = TF;            // compile-time: the plain variant carries no activation code or coefficient registers
  if (tf) {
    for (int i = threadIdx.x; i < C; i += 256) {
      ssm[i] = scale[i];
      ssm[C + i] = shift[i];
    }
  }
  __syncthreads();
  const unsigned groups = (unsigned)(C >> 3);
  const unsigned strips = (unsigned)((Q + TW - 1) / TW);
  const unsigned total = (unsigned)N * (unsigned)P * strips * groups;       // < 2^31 (checked by the launcher)
  // Each block owns a CONTIGUOUS range of items (order: image, output row, strip, channel group) = a band of consecutive
  // rows of one image, so the 3x vertical re-use of the input rows hits this SM's L1 instead of going to L2 (with a
  // grid-stride order the kernel sat at the L2 -> SM limit: 4.5 loads per input pixel).  Threads step by S = the largest
  // multiple of the group count <= 256, so a thread keeps its channel group (register-resident statistics).
  const unsigned S = 256u / groups * groups;
  unsigned chunk = (total + gridDim.x - 1) / gridDim.x;
  chunk = (chunk + S - 1) / S * S;
  const unsigned begin = blockIdx.x * chunk;
  const unsigned end = min(total, begin + chunk);
  const bool worker = threadIdx.x < S;
  const int grp = (int)(threadIdx.x % groups);                              // begin is a multiple of S, S of groups
  const float* wg = wsm + grp * 8;
  float s1[8], s2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s1[j] = s2[j] = 0.f;
  for (unsigned i = begin + threadIdx.x; worker && i < end; i += S) {
    unsigned px = i / groups;
    const int qs = (int)(px % strips);
    px /= strips;
    const int p = (int)(px % (unsigned)P);
    const long n = px / (unsigned)P;
    const int q0 = qs * TW;
    const int w0 = q0 * STRIDE - 1;
    float acc[TW][8];
#pragma unroll
    for (int t = 0; t < TW; ++t)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[t][j] = 0.f;
#pragma unroll 1          // (fully unrolled rows at one block per SM measured slower: 251 vs 217 us)
    for (int r = 0; r < 3; ++r) {
      const int h = p * STRIDE - 1 + r;
      const bool row_ok = h >= 0 && h < H;
      const bf16* xrow = x + ((n * H + (row_ok ? h : 0)) * W) * C + grp * 8;
      uint4 u[NIN];
      bool ok[NIN];
#pragma unroll
      for (int c = 0; c < NIN; ++c) {
        ok[c] = row_ok && (w0 + c) >= 0 && (w0 + c) < W;
        u[c] = ok[c] ? *reinterpret_cast<const uint4*>(xrow + (long)(w0 + c) * C) : make_uint4(0u, 0u, 0u, 0u);
      }
      float wk[3][8];
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const float4 w0v = *reinterpret_cast<const float4*>(wg + (r * 3 + s) * C);
        const float4 w1v = *reinterpret_cast<const float4*>(wg + (r * 3 + s) * C + 4);
        wk[s][0] = w0v.x; wk[s][1] = w0v.y; wk[s][2] = w0v.z; wk[s][3] = w0v.w;
        wk[s][4] = w1v.x; wk[s][5] = w1v.y; wk[s][6] = w1v.z; wk[s][7] = w1v.w;
      }
      float sc[8], sh[8];
      if (tf) {
        const float4 a0 = *reinterpret_cast<const float4*>(ssm + grp * 8), a1 = *reinterpret_cast<const float4*>(ssm + grp * 8 + 4);
        const float4 b0 = *reinterpret_cast<const float4*>(ssm + C + grp * 8), b1 = *reinterpret_cast<const float4*>(ssm + C + grp * 8 + 4);
        sc[0] = a0.x; sc[1] = a0.y; sc[2] = a0.z; sc[3] = a0.w; sc[4] = a1.x; sc[5] = a1.y; sc[6] = a1.z; sc[7] = a1.w;
        sh[0] = b0.x; sh[1] = b0.y; sh[2] = b0.z; sh[3] = b0.w; sh[4] = b1.x; sh[5] = b1.y; sh[6] = b1.z; sh[7] = b1.w;
      }
#pragma unroll
      for (int c = 0; c < NIN; ++c) {
        float a[8];
        unpack8(u[c], a);
        if (tf) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float v = fmaf(a[j], sc[j], sh[j]);
            if (act >= 1) v = fmaxf(v, 0.f);
            if (act == 2) v = fminf(v, 6.f);
            a[j] = ok[c] ? bf16_round(v) : 0.f;     // the activation is a bf16 tensor in the unfused formulation
          }
        }
#pragma unroll
        for (int t = 0; t < TW; ++t) {
          const int s = c - t * STRIDE;             // compile-time after unrolling
          if (s >= 0 && s < 3) {
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[t][j] = fmaf(a[j], wk[s][j], acc[t][j]);
          }
        }
      }
    }
    bf16* yrow = y + (((n * P + p) * Q) + q0) * C + grp * 8;
#pragma unroll
    for (int t = 0; t < TW; ++t) {
      if (q0 + t < Q) {
        const uint4 o = pack8(acc[t]);
        *reinterpret_cast<uint4*>(yrow + (long)t * C) = o;
        if (sum != nullptr) {
          float r8[8];
          unpack8(o, r8);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            s1[j] += r8[j];
            s2[j] = fmaf(r8[j], r8[j], s2[j]);
          }
        }
      }
    }
  }
  if (sum != nullptr) {
    __syncthreads();                                 // everyone is done with the filter taps: reuse the bytes
    flush_stats_local(s1, s2, (int)groups, worker, sm, sum, sumsq);
  }
}

unsigned blocks_for(long items, int groups) {
  long b = (items + 255) / 256;
  const long cap = (long)b2_num_sms() * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  if (256 % groups != 0) b = b / groups * groups > 0 ? b / groups * groups : groups;   // grid stride % groups == 0
  return (unsigned)b;
}

}  // namespace

// y [N,P,Q,32] bf16 = Conv2d(3, 32, 3, stride 2, pad 1)(x [N,3,H,W] fp32 or bf16); w [32,3,3,3] fp32 (torch layout);
// sum / sumsq [32] ACCUMULATED when given
B2_API int b2_mbv2_stem_conv(const void* x, int in_bf16, const float* w, void* y, float* sum, float* sumsq, int N, int H, int W,
                             void* stream) {
  B2_ARG_CHECK(x && w && y && N > 0 && H > 0 && W > 0, "b2_mbv2_stem_conv: null pointer or empty");
  B2_ARG_CHECK((sum == nullptr) == (sumsq == nullptr), "b2_mbv2_stem_conv: sum and sumsq go together");
  const int P = (H + 2 - 3) / 2 + 1, Q = (W + 2 - 3) / 2 + 1;
  const unsigned blocks = blocks_for((long)N * P * Q * 4, 4);
  cudaStream_t st = (cudaStream_t)stream;
  if (in_bf16)
    stem3x3s2_kernel<bf16><<<blocks, 256, 0, st>>>((const bf16*)x, w, (bf16*)y, sum, sumsq, N, H, W, P, Q);
  else
    stem3x3s2_kernel<float><<<blocks, 256, 0, st>>>((const float*)x, w, (bf16*)y, sum, sumsq, N, H, W, P, Q);
  B2_LAUNCH_CHECK("stem3x3s2_kernel");
  return 0;
}

// y [N,P,Q,C] bf16 = depthwise 3x3 (stride 1 or 2, pad 1) of act(x * scale + shift); w [C,1,3,3] fp32; statistics ACCUMULATED
B2_API int b2_dwconv3x3_bn_nhwc_bf16(const void* x, const float* scale, const float* shift, int act, const float* w, void* y,
                                     float* sum, float* sumsq, int N, int H, int W, int C, int stride, void* stream) {
  B2_ARG_CHECK(x && w && y && N > 0 && H > 0 && W > 0, "b2_dwconv3x3_bn_nhwc_bf16: null pointer or empty");
  B2_ARG_CHECK(C % 8 == 0 && C >= 8 && C <= 1024, "b2_dwconv3x3_bn_nhwc_bf16: C must be a multiple of 8 in [8, 1024]");
  B2_ARG_CHECK(stride == 1 || stride == 2, "b2_dwconv3x3_bn_nhwc_bf16: stride 1 or 2");
  B2_ARG_CHECK((scale == nullptr) == (shift == nullptr) && (sum == nullptr) == (sumsq == nullptr),
               "b2_dwconv3x3_bn_nhwc_bf16: scale/shift and sum/sumsq go in pairs");
  const int P = (H + 2 - 3) / stride + 1, Q = (W + 2 - 3) / stride + 1;
  const int groups = C / 8;
  B2_ARG_CHECK((long)N * P * Q * groups < (1L << 31), "b2_dwconv3x3_bn_nhwc_bf16: too many work items");
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned b1 = blocks_for((long)N * P * ((Q + 3) / 4) * groups, groups), b2 = blocks_for((long)N * P * ((Q + 1) / 2) * groups, groups);
#define B2_DW(S, T, B) dwconv3x3_kernel<S, T><<<B, 256, 0, st>>>((const bf16*)x, scale, shift, act, w, (bf16*)y, sum, sumsq, N, H, W, C, P, Q)
  if (stride == 1 && scale != nullptr) B2_DW(1, true, b1);
  else if (stride == 1) B2_DW(1, false, b1);
  else if (scale != nullptr) B2_DW(2, true, b2);
  else B2_DW(2, false, b2);
#undef B2_DW
  B2_LAUNCH_CHECK("dwconv3x3_kernel");
  return 0;
}
