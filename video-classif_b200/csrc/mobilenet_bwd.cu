// Backward kernels of the trainable MobileNetV2 frame encoder (lrcn/lrcn.py:196-230,246-283 and lrcn/rgb_lrcn.py:208-227 with
// CNN_BACKBONE = "mobilenet_v2": `freeze_cnn_layers` leaves the whole backbone trainable when CONF_FINETUNE is set):
//   * dwconv3x3_dgrad_kernel / dwconv3x3_wgrad_kernel   data and weight gradient of the depthwise 3x3 convs (stride 1 / 2, pad 1)
//   * stem3x3s2_wgrad_kernel                            weight gradient of Conv2d(3, 32, 3, stride 2, pad 1) from the NCHW frames
// The 1x1 convs take their weight gradient from the tcgen05 kernel of conv_bwd.cu (any C % 8 == 0 in its 1x1 form) and their data
// gradient from the tcgen05 GEMM; BatchNorm backward with the ReLU6 mask is b2_bn_bwd_relu6_nhwc_bf16 (conv_bwd.cu).
// Every tensor here is memory bound: 128-bit accesses, fp32 accumulation, one pass each.
#include "common.cuh"

namespace {

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  f[0] = __uint_as_float(u.x << 16); f[1] = __uint_as_float(u.x & 0xffff0000u);
  f[2] = __uint_as_float(u.y << 16); f[3] = __uint_as_float(u.y & 0xffff0000u);
  f[4] = __uint_as_float(u.z << 16); f[5] = __uint_as_float(u.z & 0xffff0000u);
  f[6] = __uint_as_float(u.w << 16); f[7] = __uint_as_float(u.w & 0xffff0000u);
}
__device__ __forceinline__ uint4 pack8(const float (&o)[8]) {
  return make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
}

// data gradient: dx[n, iy, ix, c] = sum_{r,s} dy[n, (iy + 1 - r) / S, (ix + 1 - s) / S, c] * w[c][r][s] over the taps whose
// output position exists; thread = (input pixel, 8 channels)
template <int S>
__global__ void __launch_bounds__(256)
dwconv3x3_dgrad_kernel(const bf16* __restrict__ dy, const float* __restrict__ w, bf16* __restrict__ dx, int N, int H, int W, int C,
                       int P, int Q) {
  const int groups = C >> 3;
  const long total = (long)N * H * W * groups;
  const long stride = (long)gridDim.x * blockDim.x;              // a multiple of `groups` (host): fixed channel group per thread
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int grp = (int)(i % groups);
  float wk[9][8];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int j = 0; j < 8; ++j) wk[t][j] = w[(grp * 8 + j) * 9 + t];
  for (; i < total; i += stride) {
    const unsigned pix = (unsigned)(i / groups);
    const unsigned t1 = pix / (unsigned)W;
    const int ix = (int)(pix - t1 * (unsigned)W);
    const unsigned n = t1 / (unsigned)H;
    const int iy = (int)(t1 - n * (unsigned)H);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int ny = iy + 1 - r;
      if (ny < 0 || (S == 2 && (ny & 1)) || ny / S >= P) continue;
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const int nx = ix + 1 - s;
        if (nx < 0 || (S == 2 && (nx & 1)) || nx / S >= Q) continue;
        float g[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(dy + (((long)n * P + ny / S) * Q + nx / S) * C + grp * 8)), g);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(g[j], wk[r * 3 + s][j], acc[j]);
      }
    }
    *reinterpret_cast<uint4*>(dx + (long)pix * C + grp * 8) = pack8(acc);
  }
}

// weight gradient: dw[c][r][s] (fp32, ACCUMULATED) = sum_{n,p,q} dy[n,p,q,c] * x[n, p S + r - 1, q S + s - 1, c];
// thread = (output pixel, 8 channels) with 72 register accumulators over its grid-stride walk, combined through shared memory
template <int S>
__global__ void __launch_bounds__(256)
dwconv3x3_wgrad_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dy, float* __restrict__ dw, int N, int H, int W, int C,
                       int P, int Q) {
  extern __shared__ float dw_s[];                                 // [C][9]
  const int groups = C >> 3;
  for (int k = threadIdx.x; k < C * 9; k += blockDim.x) dw_s[k] = 0.f;
  __syncthreads();
  const long total = (long)N * P * Q * groups;
  const long stride = (long)gridDim.x * blockDim.x;
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int grp = (int)(i % groups);
  float acc[9][8];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[t][j] = 0.f;
  for (; i < total; i += stride) {
    const unsigned pix = (unsigned)(i / groups);
    const unsigned t1 = pix / (unsigned)Q;
    const int q = (int)(pix - t1 * (unsigned)Q);
    const unsigned n = t1 / (unsigned)P;
    const int p = (int)(t1 - n * (unsigned)P);
    float g[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(dy + (long)pix * C + grp * 8)), g);
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int iy = p * S + r - 1;
      if (iy < 0 || iy >= H) continue;
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const int ix = q * S + s - 1;
        if (ix < 0 || ix >= W) continue;
        float xv[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(x + (((long)n * H + iy) * W + ix) * C + grp * 8)), xv);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[r * 3 + s][j] = fmaf(g[j], xv[j], acc[r * 3 + s][j]);
      }
    }
  }
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(&dw_s[(grp * 8 + j) * 9 + t], acc[t][j]);
  __syncthreads();
  for (int k = threadIdx.x; k < C * 9; k += blockDim.x) atomicAdd(dw + k, dw_s[k]);
}

// weight gradient of the stem Conv2d(3, 32, 3, stride 2, pad 1): dw[co][ci][r][s] (fp32, torch layout, ACCUMULATED) from the NCHW
// frames and dy bf16 NHWC [N,P,Q,32].  CTA = 16 x 16 output pixels; warp = 32 of them; lane = 4 output channels x 7 (ci, tap) pairs
template <typename InT>
__global__ void __launch_bounds__(256)
stem3x3s2_wgrad_kernel(const InT* __restrict__ x, const bf16* __restrict__ dy, float* __restrict__ dw, int N, int H, int W, int P,
                       int Q) {
  __shared__ float xs[3][33][33];
  __shared__ __align__(16) float ds[256][32];
  const int tiles_x = (Q + 15) >> 4, tiles_y = (P + 15) >> 4;
  const long items = (long)N * tiles_x * tiles_y;
  const int pg = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cob = lane & 7, jb = lane >> 3;                        // channels 4 cob .. ; pairs 7 jb .. 7 jb + 6 (of 27)
  int xoff[7];
#pragma unroll
  for (int k = 0; k < 7; ++k) {
    const int j = min(jb * 7 + k, 26);
    const int ci = j / 9, tap = j - ci * 9;
    xoff[k] = (ci * 33 + tap / 3) * 33 + tap % 3;
  }
  float acc[4][7];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int k = 0; k < 7; ++k) acc[a][k] = 0.f;
  const float* xsf = &xs[0][0][0];
  for (long it = blockIdx.x; it < items; it += gridDim.x) {
    const int n = (int)(it / (tiles_x * tiles_y));
    const int tr = (int)(it - (long)n * tiles_x * tiles_y);
    const int q0 = (tr % tiles_x) << 4, p0 = (tr / tiles_x) << 4;
    __syncthreads();
    for (int idx = threadIdx.x; idx < 3 * 33 * 33; idx += 256) {
      const int ci = idx / 1089, rem = idx - ci * 1089;
      const int yy = rem / 33, xx = rem - yy * 33;
      const int gy = 2 * p0 + yy - 1, gx = 2 * q0 + xx - 1;
      xs[ci][yy][xx] = (gy >= 0 && gy < H && gx >= 0 && gx < W) ? (float)x[(((long)n * 3 + ci) * H + gy) * W + gx] : 0.f;
    }
    {
      const int px = threadIdx.x & 15, py = threadIdx.x >> 4;
      const int p = p0 + py, q = q0 + px;
      const bool ok = p < P && q < Q;
      const uint4* src = reinterpret_cast<const uint4*>(dy + (((long)n * P + p) * Q + q) * 32);
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        float g[8];
        unpack8(ok ? __ldg(src + v) : make_uint4(0u, 0u, 0u, 0u), g);
#pragma unroll
        for (int j = 0; j < 8; ++j) ds[threadIdx.x][v * 8 + j] = g[j];
      }
    }
    __syncthreads();
#pragma unroll 2
    for (int i = 0; i < 32; ++i) {
      const int pl = pg * 32 + i;                       // pixel (py = pl >> 4, px = pl & 15) -> input window origin (2 py, 2 px)
      const float4 d4 = *reinterpret_cast<const float4*>(&ds[pl][cob * 4]);
      const int base = (2 * (pl >> 4)) * 33 + 2 * (pl & 15);
#pragma unroll
      for (int k = 0; k < 7; ++k) {
        const float xv = xsf[base + xoff[k]];
        acc[0][k] = fmaf(d4.x, xv, acc[0][k]);
        acc[1][k] = fmaf(d4.y, xv, acc[1][k]);
        acc[2][k] = fmaf(d4.z, xv, acc[2][k]);
        acc[3][k] = fmaf(d4.w, xv, acc[3][k]);
      }
    }
  }
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int k = 0; k < 7; ++k) {
      const int j = jb * 7 + k;
      if (j < 27) atomicAdd(dw + (cob * 4 + a) * 27 + j, acc[a][k]);
    }
}

// grid with a stride (blocks * 256) that is a multiple of the channel-group count, so a thread's channel group is fixed
unsigned blocks_for_groups(long items, int groups, long cap) {
  long b = (items + 255) / 256;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  if (256 % groups != 0) b = b / groups * groups > 0 ? b / groups * groups : groups;
  return (unsigned)b;
}

}  // namespace

// dx [N,H,W,C] bf16 = data gradient of the depthwise 3x3 conv (stride 1 or 2, pad 1) w [C,1,3,3] fp32 given dy [N,P,Q,C]
B2_API int b2_dwconv3x3_dgrad_nhwc_bf16(const void* dy, const float* w, void* dx, int N, int H, int W, int C, int stride, void* stream) {
  B2_ARG_CHECK(dy && w && dx && N > 0 && H > 0 && W > 0, "b2_dwconv3x3_dgrad_nhwc_bf16: null pointer or empty");
  B2_ARG_CHECK(C % 8 == 0 && C >= 8 && C <= 2048 && (stride == 1 || stride == 2), "b2_dwconv3x3_dgrad_nhwc_bf16: C % 8, stride 1 / 2");
  const int P = (H + 2 - 3) / stride + 1, Q = (W + 2 - 3) / stride + 1;
  const int groups = C / 8;
  B2_ARG_CHECK((long)N * H * W < (1L << 31), "b2_dwconv3x3_dgrad_nhwc_bf16: too many pixels");
  const unsigned blocks = blocks_for_groups((long)N * H * W * groups, groups, (long)b2_num_sms() * 16);
  if (stride == 1) dwconv3x3_dgrad_kernel<1><<<blocks, 256, 0, (cudaStream_t)stream>>>((const bf16*)dy, w, (bf16*)dx, N, H, W, C, P, Q);
  else dwconv3x3_dgrad_kernel<2><<<blocks, 256, 0, (cudaStream_t)stream>>>((const bf16*)dy, w, (bf16*)dx, N, H, W, C, P, Q);
  B2_LAUNCH_CHECK("dwconv3x3_dgrad_kernel");
  return 0;
}

// dw [C,1,3,3] fp32 (ACCUMULATED, caller zeroes) = weight gradient of the depthwise conv from x [N,H,W,C] and dy [N,P,Q,C]
B2_API int b2_dwconv3x3_wgrad_nhwc_bf16(const void* x, const void* dy, float* dw, int N, int H, int W, int C, int stride, void* stream) {
  B2_ARG_CHECK(x && dy && dw && N > 0 && H > 0 && W > 0, "b2_dwconv3x3_wgrad_nhwc_bf16: null pointer or empty");
  B2_ARG_CHECK(C % 8 == 0 && C >= 8 && C <= 1024 && (stride == 1 || stride == 2),
               "b2_dwconv3x3_wgrad_nhwc_bf16: C % 8 in [8, 1024], stride 1 / 2");
  const int P = (H + 2 - 3) / stride + 1, Q = (W + 2 - 3) / stride + 1;
  const int groups = C / 8;
  B2_ARG_CHECK((long)N * P * Q < (1L << 31), "b2_dwconv3x3_wgrad_nhwc_bf16: too many pixels");
  // few, long-lived CTAs: one shared-memory flush each
  const unsigned blocks = blocks_for_groups((long)N * P * Q * groups, groups, (long)b2_num_sms() * 2);
  const size_t smem = (size_t)C * 9 * sizeof(float);
  if (stride == 1)
    dwconv3x3_wgrad_kernel<1><<<blocks, 256, smem, (cudaStream_t)stream>>>((const bf16*)x, (const bf16*)dy, dw, N, H, W, C, P, Q);
  else
    dwconv3x3_wgrad_kernel<2><<<blocks, 256, smem, (cudaStream_t)stream>>>((const bf16*)x, (const bf16*)dy, dw, N, H, W, C, P, Q);
  B2_LAUNCH_CHECK("dwconv3x3_wgrad_kernel");
  return 0;
}

// dw [32,3,3,3] fp32 (ACCUMULATED) = weight gradient of the stem conv from the NCHW frames and dy [N,P,Q,32] bf16
B2_API int b2_mbv2_stem_wgrad(const void* x, int in_bf16, const void* dy, float* dw, int N, int H, int W, void* stream) {
  B2_ARG_CHECK(x && dy && dw && N > 0 && H > 0 && W > 0, "b2_mbv2_stem_wgrad: null pointer or empty");
  const int P = (H + 2 - 3) / 2 + 1, Q = (W + 2 - 3) / 2 + 1;
  const long items = (long)N * ((Q + 15) / 16) * ((P + 15) / 16);
  const long cap = (long)b2_num_sms() * 2;
  const unsigned blocks = (unsigned)(items < cap ? items : cap);
  if (in_bf16)
    stem3x3s2_wgrad_kernel<bf16><<<blocks, 256, 0, (cudaStream_t)stream>>>((const bf16*)x, (const bf16*)dy, dw, N, H, W, P, Q);
  else
    stem3x3s2_wgrad_kernel<float><<<blocks, 256, 0, (cudaStream_t)stream>>>((const float*)x, (const bf16*)dy, dw, N, H, W, P, Q);
  B2_LAUNCH_CHECK("stem3x3s2_wgrad_kernel");
  return 0;
}
