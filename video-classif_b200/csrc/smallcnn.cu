// fp32 NCHW kernels of the small TimeDistributed CNN (notebook LRCN nb:148-193 /
// lrcn/backup_ucf50.py:105-151): 3x3 pad-1 convolution forward / data-grad / weight-grad,
// train-mode BatchNorm2d statistics, fused BN + ReLU (+ 2x2 max-pool) forward and backward.
// This is the fp32 parity path (max rel err 1e-4 vs the reference); channel counts are 3/16/32/64,
// far too small for tensor-core tiles to pay, so these are SIMT kernels tiled through shared memory.
#include "common.cuh"

namespace {

constexpr int CO_T = 16;  // output channels per block
constexpr int CI_T = 8;   // input channels per smem slab

// y[n,o,:,:] = bias[o] + sum_{i,r,s} x[n,i,y+r-1,x+s-1] * Wk(o,i,r,s)
// transposed=0: Wk(o,i,r,s) = w[o][i][r][s]            (forward;  w is [Cout][Cin][3][3])
// transposed=1: Wk(o,i,r,s) = w[i][o][2-r][2-s]        (data grad; w is [Cin_k][Cout_k][3][3])
__global__ void __launch_bounds__(256)
conv3x3_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
               float* __restrict__ y, int N, int Cin, int Cout, int H, int W, int transposed) {
  __shared__ float xs[CI_T][18][18];
  __shared__ __align__(16) float ws[CI_T][9][CO_T];
  const int tiles_x = (W + 15) >> 4;
  const int tx0 = (blockIdx.x % tiles_x) << 4, ty0 = (blockIdx.x / tiles_x) << 4;
  const int co0 = blockIdx.y * CO_T;
  const int n = blockIdx.z;
  const int lx = threadIdx.x & 15, ly = threadIdx.x >> 4;
  float acc[CO_T];
#pragma unroll
  for (int o = 0; o < CO_T; ++o) acc[o] = 0.f;
  for (int ci0 = 0; ci0 < Cin; ci0 += CI_T) {
    for (int idx = threadIdx.x; idx < CI_T * 324; idx += 256) {
      const int ci = idx / 324, rem = idx - ci * 324;
      const int yy = rem / 18, xx = rem - yy * 18;
      const int gy = ty0 + yy - 1, gx = tx0 + xx - 1;
      float v = 0.f;
      if (ci0 + ci < Cin && gy >= 0 && gy < H && gx >= 0 && gx < W)
        v = x[(((long)n * Cin + ci0 + ci) * H + gy) * W + gx];
      xs[ci][yy][xx] = v;
    }
    for (int idx = threadIdx.x; idx < CI_T * 9 * CO_T; idx += 256) {
      const int co = idx % CO_T, tap = (idx / CO_T) % 9, ci = idx / (CO_T * 9);
      const int o = co0 + co, i = ci0 + ci;
      float v = 0.f;
      if (o < Cout && i < Cin) v = transposed ? w[((long)i * Cout + o) * 9 + (8 - tap)] : w[((long)o * Cin + i) * 9 + tap];
      ws[ci][tap][co] = v;
    }
    __syncthreads();
#pragma unroll
    for (int ci = 0; ci < CI_T; ++ci) {
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const float v = xs[ci][ly + tap / 3][lx + tap % 3];
        const float4* wp = reinterpret_cast<const float4*>(&ws[ci][tap][0]);
#pragma unroll
        for (int q = 0; q < CO_T / 4; ++q) {
          const float4 wv = wp[q];
          acc[q * 4 + 0] = fmaf(v, wv.x, acc[q * 4 + 0]);
          acc[q * 4 + 1] = fmaf(v, wv.y, acc[q * 4 + 1]);
          acc[q * 4 + 2] = fmaf(v, wv.z, acc[q * 4 + 2]);
          acc[q * 4 + 3] = fmaf(v, wv.w, acc[q * 4 + 3]);
        }
      }
    }
    __syncthreads();
  }
  const int gy = ty0 + ly, gx = tx0 + lx;
  if (gy < H && gx < W) {
#pragma unroll
    for (int o = 0; o < CO_T; ++o)
      if (co0 + o < Cout)
        y[(((long)n * Cout + co0 + o) * H + gy) * W + gx] = acc[o] + (bias ? bias[co0 + o] : 0.f);
  }
}

// dw[o][i][r][s] += sum_{n,y,x} dy[n,o,y,x] * x[n,i,y+r-1,x+s-1]
// block: CO_T output channels x all input channels, loops over (n, 16x16 tile) work items and keeps
// its partial dw in registers; thread = (co, input-channel group); sliding 3x3 window along x.
constexpr int WG_CI_GROUPS = 16;
template <int CI_PER>  // input channels per thread (Cin <= 16*CI_PER)
__global__ void __launch_bounds__(256)
conv3x3_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dw, int N, int Cin,
                     int Cout, int H, int W) {
  extern __shared__ float sm[];
  const int cin_pad = WG_CI_GROUPS * CI_PER;
  float* xs = sm;                         // [cin_pad][18][19]
  float* ds = sm + cin_pad * 18 * 19;     // [16][16][CO_T]
  const int co = threadIdx.x % CO_T;
  const int cg = threadIdx.x / CO_T;      // 0..15
  const int co0 = blockIdx.y * CO_T;
  const int tiles_x = (W + 15) >> 4, tiles_y = (H + 15) >> 4;
  const long items = (long)N * tiles_x * tiles_y;
  float acc[CI_PER][9];
#pragma unroll
  for (int c = 0; c < CI_PER; ++c)
#pragma unroll
    for (int t = 0; t < 9; ++t) acc[c][t] = 0.f;
  for (long it = blockIdx.x; it < items; it += gridDim.x) {
    const int n = (int)(it / (tiles_x * tiles_y));
    const int tr = (int)(it % (tiles_x * tiles_y));
    const int tx0 = (tr % tiles_x) << 4, ty0 = (tr / tiles_x) << 4;
    for (int idx = threadIdx.x; idx < cin_pad * 324; idx += 256) {
      const int ci = idx / 324, rem = idx - ci * 324;
      const int yy = rem / 18, xx = rem - yy * 18;
      const int gy = ty0 + yy - 1, gx = tx0 + xx - 1;
      float v = 0.f;
      if (ci < Cin && gy >= 0 && gy < H && gx >= 0 && gx < W) v = x[(((long)n * Cin + ci) * H + gy) * W + gx];
      xs[(ci * 18 + yy) * 19 + xx] = v;
    }
    for (int idx = threadIdx.x; idx < 256 * CO_T; idx += 256) {
      const int px = idx & 15, py = (idx >> 4) & 15, o = idx >> 8;
      const int gy = ty0 + py, gx = tx0 + px;
      float v = 0.f;
      if (co0 + o < Cout && gy < H && gx < W) v = dy[(((long)n * Cout + co0 + o) * H + gy) * W + gx];
      ds[(py * 16 + px) * CO_T + o] = v;
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < CI_PER; ++c) {
      const int ci = cg * CI_PER + c;
      const float* xc = xs + ci * 18 * 19;
      for (int py = 0; py < 16; ++py) {
        float w0[3], w1[3], w2[3];  // window columns: rows py..py+2
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          w0[r] = xc[(py + r) * 19 + 0];
          w1[r] = xc[(py + r) * 19 + 1];
        }
#pragma unroll 4
        for (int px = 0; px < 16; ++px) {
#pragma unroll
          for (int r = 0; r < 3; ++r) w2[r] = xc[(py + r) * 19 + px + 2];
          const float d = ds[(py * 16 + px) * CO_T + co];
#pragma unroll
          for (int r = 0; r < 3; ++r) {
            acc[c][r * 3 + 0] = fmaf(d, w0[r], acc[c][r * 3 + 0]);
            acc[c][r * 3 + 1] = fmaf(d, w1[r], acc[c][r * 3 + 1]);
            acc[c][r * 3 + 2] = fmaf(d, w2[r], acc[c][r * 3 + 2]);
          }
#pragma unroll
          for (int r = 0; r < 3; ++r) {
            w0[r] = w1[r];
            w1[r] = w2[r];
          }
        }
      }
    }
    __syncthreads();
  }
  if (co0 + co < Cout) {
#pragma unroll
    for (int c = 0; c < CI_PER; ++c) {
      const int ci = cg * CI_PER + c;
      if (ci < Cin) {
#pragma unroll
        for (int t = 0; t < 9; ++t) atomicAdd(dw + (((long)(co0 + co) * Cin + ci) * 9 + t), acc[c][t]);
      }
    }
  }
}

// per-channel sum / sum of squares over (N, H*W) of an NCHW tensor; double accumulators
__global__ void __launch_bounds__(256)
bn2d_stats_kernel(const float* __restrict__ x, int N, int C, int HW, double* __restrict__ sum,
                  double* __restrict__ sumsq) {
  __shared__ double red1[8], red2[8];
  const int c = blockIdx.x;
  double s1 = 0.0, s2 = 0.0;
  for (int n = blockIdx.y; n < N; n += gridDim.y) {
    const float* p = x + ((long)n * C + c) * HW;
    for (int i = threadIdx.x; i < HW; i += 256) {
      const float v = p[i];
      s1 += v;
      s2 += (double)v * v;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  if ((threadIdx.x & 31) == 0) {
    red1[threadIdx.x >> 5] = s1;
    red2[threadIdx.x >> 5] = s2;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0, b = 0;
    for (int i = 0; i < 8; ++i) {
      a += red1[i];
      b += red2[i];
    }
    atomicAdd(sum + c, a);
    if (sumsq) atomicAdd(sumsq + c, b);
  }
}

// scale/shift/mean/rstd from the statistics; running-stat update (momentum, unbiased variance)
__global__ void bn2d_finalize_kernel(const double* __restrict__ sum, const double* __restrict__ sumsq, long count,
                                     const float* __restrict__ gamma, const float* __restrict__ beta,
                                     float* __restrict__ running_mean, float* __restrict__ running_var, float momentum,
                                     float eps, int train, float* __restrict__ scale, float* __restrict__ shift,
                                     float* __restrict__ mean_out, float* __restrict__ rstd_out, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float mean, var;
  if (train) {
    const double m = sum[c] / (double)count;
    double v = sumsq[c] / (double)count - m * m;
    if (v < 0) v = 0;
    mean = (float)m;
    var = (float)v;
    if (running_mean) {
      const double unb = count > 1 ? v * (double)count / (double)(count - 1) : v;
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unb;
    }
  } else {
    mean = running_mean[c];
    var = running_var[c];
  }
  const float rstd = 1.0f / sqrtf(var + eps);
  const float sc = gamma[c] * rstd;
  scale[c] = sc;
  shift[c] = beta[c] - mean * sc;
  mean_out[c] = mean;
  rstd_out[c] = rstd;
}

// y = [maxpool2x2] relu(x*scale+shift)
__global__ void __launch_bounds__(256)
bn2d_act_pool_fwd_kernel(const float* __restrict__ x, const float* __restrict__ scale, const float* __restrict__ shift,
                         float* __restrict__ y, bf16* __restrict__ y_bf16, int N, int C, int H, int W, int pool) {
  const int Ho = pool ? H >> 1 : H, Wo = pool ? W >> 1 : W;
  const long total = (long)N * C * Ho * Wo;
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int xo = (int)(idx % Wo);
    const int yo = (int)((idx / Wo) % Ho);
    const long nc = idx / ((long)Wo * Ho);
    const int c = (int)(nc % C);
    const float sc = scale[c], sh = shift[c];
    const float* p = x + nc * H * W;
    float o;
    if (pool) {
      const float* q = p + (long)(2 * yo) * W + 2 * xo;
      const float a0 = fmaxf(fmaf(q[0], sc, sh), 0.f), a1 = fmaxf(fmaf(q[1], sc, sh), 0.f);
      const float a2 = fmaxf(fmaf(q[W], sc, sh), 0.f), a3 = fmaxf(fmaf(q[W + 1], sc, sh), 0.f);
      o = fmaxf(fmaxf(a0, a1), fmaxf(a2, a3));
    } else {
      o = fmaxf(fmaf(p[(long)yo * W + xo], sc, sh), 0.f);
    }
    y[idx] = o;
    if (y_bf16) y_bf16[idx] = __float2bfloat16_rn(o);
  }
}

// gradient wrt the BN output z at input position (yy,xx) given the pooled/activated upstream grad
__device__ __forceinline__ void pool_relu_grad(const float* __restrict__ q, int W, float sc, float sh, float g, int pool,
                                               float (&dz)[4], float (&xv)[4]) {
  if (pool) {
    xv[0] = q[0];
    xv[1] = q[1];
    xv[2] = q[W];
    xv[3] = q[W + 1];
    float a[4];
    int best = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = fmaxf(fmaf(xv[i], sc, sh), 0.f);
#pragma unroll
    for (int i = 1; i < 4; ++i)
      if (a[i] > a[best]) best = i;   // first maximum wins, like torch max_pool2d
#pragma unroll
    for (int i = 0; i < 4; ++i) dz[i] = (i == best && a[i] > 0.f) ? g : 0.f;
  } else {
    xv[0] = q[0];
    dz[0] = fmaf(xv[0], sc, sh) > 0.f ? g : 0.f;
  }
}

// pass 1: s1[c] = sum dz, s2[c] = sum dz * xhat
__global__ void __launch_bounds__(256)
bn2d_bwd_reduce_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ scale,
                       const float* __restrict__ shift, const float* __restrict__ mean, const float* __restrict__ rstd,
                       int N, int C, int H, int W, int pool, double* __restrict__ s1, double* __restrict__ s2) {
  __shared__ double red1[8], red2[8];
  const int c = blockIdx.x;
  const int Ho = pool ? H >> 1 : H, Wo = pool ? W >> 1 : W;
  const float sc = scale[c], sh = shift[c], mu = mean[c], rs = rstd[c];
  double a1 = 0, a2 = 0;
  for (int n = blockIdx.y; n < N; n += gridDim.y) {
    const float* px = x + ((long)n * C + c) * H * W;
    const float* pd = dy + ((long)n * C + c) * Ho * Wo;
    for (int i = threadIdx.x; i < Ho * Wo; i += 256) {
      const int yo = i / Wo, xo = i - yo * Wo;
      const float* q = pool ? px + (long)(2 * yo) * W + 2 * xo : px + i;
      float dz[4], xv[4];
      pool_relu_grad(q, W, sc, sh, pd[i], pool, dz, xv);
      const int cnt = pool ? 4 : 1;
      for (int k = 0; k < cnt; ++k) {
        a1 += dz[k];
        a2 += dz[k] * ((xv[k] - mu) * rs);
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a1 += __shfl_xor_sync(0xffffffffu, a1, o);
    a2 += __shfl_xor_sync(0xffffffffu, a2, o);
  }
  if ((threadIdx.x & 31) == 0) {
    red1[threadIdx.x >> 5] = a1;
    red2[threadIdx.x >> 5] = a2;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double u = 0, v = 0;
    for (int i = 0; i < 8; ++i) {
      u += red1[i];
      v += red2[i];
    }
    atomicAdd(s1 + c, u);
    atomicAdd(s2 + c, v);
  }
}

// pass 2: dx = gamma*rstd*(dz - s1/n - xhat*s2/n)   (train)   or   gamma*rstd*dz   (eval)
__global__ void __launch_bounds__(256)
bn2d_bwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ scale,
                      const float* __restrict__ shift, const float* __restrict__ mean, const float* __restrict__ rstd,
                      const float* __restrict__ gamma, const double* __restrict__ s1, const double* __restrict__ s2,
                      long count, int train, float* __restrict__ dx, int N, int C, int H, int W, int pool) {
  const int Ho = pool ? H >> 1 : H, Wo = pool ? W >> 1 : W;
  const long total = (long)N * C * Ho * Wo;
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int xo = (int)(idx % Wo);
    const int yo = (int)((idx / Wo) % Ho);
    const long nc = idx / ((long)Wo * Ho);
    const int c = (int)(nc % C);
    const float sc = scale[c], sh = shift[c], mu = mean[c], rs = rstd[c];
    const float m1 = train ? (float)(s1[c] / (double)count) : 0.f;
    const float m2 = train ? (float)(s2[c] / (double)count) : 0.f;
    const float gr = gamma[c] * rs;
    const long off = pool ? (long)(2 * yo) * W + 2 * xo : (long)yo * W + xo;
    const float* q = x + nc * H * W + off;
    float* o = dx + nc * H * W + off;
    float dz[4], xv[4];
    pool_relu_grad(q, W, sc, sh, dy[idx], pool, dz, xv);
    if (pool) {
      o[0] = gr * (dz[0] - m1 - (xv[0] - mu) * rs * m2);
      o[1] = gr * (dz[1] - m1 - (xv[1] - mu) * rs * m2);
      o[W] = gr * (dz[2] - m1 - (xv[2] - mu) * rs * m2);
      o[W + 1] = gr * (dz[3] - m1 - (xv[3] - mu) * rs * m2);
    } else {
      o[0] = gr * (dz[0] - m1 - (xv[0] - mu) * rs * m2);
    }
  }
}

__global__ void double_to_float_kernel(const double* __restrict__ s, float* __restrict__ d, int n, int accumulate) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) d[i] = accumulate ? d[i] + (float)s[i] : (float)s[i];
}

int ew_grid(long total) {
  long b = (total + 255) / 256;
  long cap = (long)b2_num_sms() * 8;
  return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

}  // namespace

B2_API int b2_conv3x3_f32(const float* x, const float* w, const float* bias, float* y, int N, int Cin, int Cout, int H,
                          int W, int transposed, void* stream) {
  B2_ARG_CHECK(x && w && y && N > 0 && Cin > 0 && Cout > 0 && H > 0 && W > 0, "b2_conv3x3_f32: null pointer or empty");
  B2_ARG_CHECK(N <= 65535, "b2_conv3x3_f32: N too large for grid.z (%d)", N);
  dim3 grid(((W + 15) / 16) * ((H + 15) / 16), b2_ceil_div(Cout, CO_T), N);
  conv3x3_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, w, bias, y, N, Cin, Cout, H, W, transposed);
  B2_LAUNCH_CHECK("conv3x3_kernel");
  return 0;
}

// dw is ACCUMULATED into (caller zeroes it)
B2_API int b2_conv3x3_wgrad_f32(const float* x, const float* dy, float* dw, int N, int Cin, int Cout, int H, int W,
                                void* stream) {
  B2_ARG_CHECK(x && dy && dw && N > 0 && Cin > 0 && Cout > 0 && H > 0 && W > 0, "b2_conv3x3_wgrad_f32: null pointer or empty");
  B2_ARG_CHECK(Cin <= 64, "b2_conv3x3_wgrad_f32: Cin=%d > 64 unsupported", Cin);
  const long items = (long)N * ((W + 15) / 16) * ((H + 15) / 16);
  const int gy = b2_ceil_div(Cout, CO_T);
  long gx = (long)b2_num_sms() * 2 / gy;
  if (gx < 1) gx = 1;
  if (gx > items) gx = items;
  dim3 grid((unsigned)gx, gy);
  cudaStream_t st = (cudaStream_t)stream;
  const int ci_per = Cin <= 16 ? 1 : (Cin <= 32 ? 2 : 4);
  const size_t smem = (size_t)(WG_CI_GROUPS * ci_per * 18 * 19 + 256 * CO_T) * sizeof(float);
  if (ci_per == 1) {
    conv3x3_wgrad_kernel<1><<<grid, 256, smem, st>>>(x, dy, dw, N, Cin, Cout, H, W);
  } else if (ci_per == 2) {
    static B2PerDeviceOnce a2;
    if (a2.needed()) { B2_CUDA_CHECK(cudaFuncSetAttribute(conv3x3_wgrad_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); a2.mark(); }
    conv3x3_wgrad_kernel<2><<<grid, 256, smem, st>>>(x, dy, dw, N, Cin, Cout, H, W);
  } else {
    static B2PerDeviceOnce a4;
    if (a4.needed()) { B2_CUDA_CHECK(cudaFuncSetAttribute(conv3x3_wgrad_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); a4.mark(); }
    conv3x3_wgrad_kernel<4><<<grid, 256, smem, st>>>(x, dy, dw, N, Cin, Cout, H, W);
  }
  B2_LAUNCH_CHECK("conv3x3_wgrad_kernel");
  return 0;
}

// sum / sumsq are ACCUMULATED into (caller zeroes them); sumsq may be null
B2_API int b2_bn2d_stats_f32(const float* x, int N, int C, int HW, double* sum, double* sumsq, void* stream) {
  B2_ARG_CHECK(x && sum && N > 0 && C > 0 && HW > 0, "b2_bn2d_stats_f32: null pointer or empty");
  int gy = (b2_num_sms() * 4 + C - 1) / C;
  if (gy > N) gy = N;
  if (gy < 1) gy = 1;
  bn2d_stats_kernel<<<dim3(C, gy), 256, 0, (cudaStream_t)stream>>>(x, N, C, HW, sum, sumsq);
  B2_LAUNCH_CHECK("bn2d_stats_kernel");
  return 0;
}

B2_API int b2_bn2d_finalize(const double* sum, const double* sumsq, long count, const float* gamma, const float* beta,
                            float* running_mean, float* running_var, float momentum, float eps, int train,
                            float* scale, float* shift, float* mean, float* rstd, int C, void* stream) {
  B2_ARG_CHECK(gamma && beta && scale && shift && mean && rstd && C > 0, "b2_bn2d_finalize: null pointer or empty");
  B2_ARG_CHECK(train ? (sum && sumsq && count > 0) : (running_mean && running_var), "b2_bn2d_finalize: missing statistics");
  bn2d_finalize_kernel<<<b2_ceil_div(C, 128), 128, 0, (cudaStream_t)stream>>>(sum, sumsq, count, gamma, beta,
                                                                              running_mean, running_var, momentum, eps,
                                                                              train, scale, shift, mean, rstd, C);
  B2_LAUNCH_CHECK("bn2d_finalize_kernel");
  return 0;
}

B2_API int b2_bn2d_act_pool_fwd_f32(const float* x, const float* scale, const float* shift, float* y, void* y_bf16,
                                    int N, int C, int H, int W, int pool, void* stream) {
  B2_ARG_CHECK(x && scale && shift && y && N > 0 && C > 0, "b2_bn2d_act_pool_fwd_f32: null pointer or empty");
  B2_ARG_CHECK(!pool || (H % 2 == 0 && W % 2 == 0), "b2_bn2d_act_pool_fwd_f32: pooling needs even H, W");
  const long total = (long)N * C * (pool ? H / 2 : H) * (pool ? W / 2 : W);
  bn2d_act_pool_fwd_kernel<<<ew_grid(total), 256, 0, (cudaStream_t)stream>>>(x, scale, shift, y, (bf16*)y_bf16, N, C, H,
                                                                             W, pool);
  B2_LAUNCH_CHECK("bn2d_act_pool_fwd_kernel");
  return 0;
}

// s1/s2 ACCUMULATED into (caller zeroes)
B2_API int b2_bn2d_act_pool_bwd_reduce_f32(const float* x, const float* dy, const float* scale, const float* shift,
                                           const float* mean, const float* rstd, int N, int C, int H, int W, int pool,
                                           double* s1, double* s2, void* stream) {
  B2_ARG_CHECK(x && dy && scale && shift && mean && rstd && s1 && s2, "b2_bn2d_act_pool_bwd_reduce_f32: null pointer");
  int gy = (b2_num_sms() * 4 + C - 1) / C;
  if (gy > N) gy = N;
  if (gy < 1) gy = 1;
  bn2d_bwd_reduce_kernel<<<dim3(C, gy), 256, 0, (cudaStream_t)stream>>>(x, dy, scale, shift, mean, rstd, N, C, H, W,
                                                                        pool, s1, s2);
  B2_LAUNCH_CHECK("bn2d_bwd_reduce_kernel");
  return 0;
}

B2_API int b2_bn2d_act_pool_bwd_apply_f32(const float* x, const float* dy, const float* scale, const float* shift,
                                          const float* mean, const float* rstd, const float* gamma, const double* s1,
                                          const double* s2, long count, int train, float* dx, int N, int C, int H,
                                          int W, int pool, void* stream) {
  B2_ARG_CHECK(x && dy && scale && shift && mean && rstd && gamma && s1 && s2 && dx, "b2_bn2d_act_pool_bwd_apply_f32: null pointer");
  const long total = (long)N * C * (pool ? H / 2 : H) * (pool ? W / 2 : W);
  bn2d_bwd_apply_kernel<<<ew_grid(total), 256, 0, (cudaStream_t)stream>>>(x, dy, scale, shift, mean, rstd, gamma, s1, s2,
                                                                          count, train, dx, N, C, H, W, pool);
  B2_LAUNCH_CHECK("bn2d_bwd_apply_kernel");
  return 0;
}

B2_API int b2_f64_to_f32(const double* src, float* dst, int n, int accumulate, void* stream) {
  B2_ARG_CHECK(src && dst && n > 0, "b2_f64_to_f32: null pointer or empty");
  double_to_float_kernel<<<b2_ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(src, dst, n, accumulate);
  B2_LAUNCH_CHECK("double_to_float_kernel");
  return 0;
}
