// Trainable-tail kernels (adapt / head layers): exact-erf GELU + LayerNorm forward/backward,
// fp32 SIMT GEMM (all transpose combinations; the fp32 parity path and the tiny-N layers),
// column sums (bias gradients), casts and cast-transposes that feed the tcgen05 GEMM.
//
// Reference semantics: medsos_lrcn/src/models.py:200-202,221-226 -- `LN(gelu(Linear(x)))` where the
// attributes named bn* are nn.LayerNorm (biased variance, eps 1e-5) and F.gelu is the erf form.
#include "common.cuh"

namespace {

constexpr float kInvSqrt2 = 0.70710678118654752440f;
constexpr float kInvSqrt2Pi = 0.39894228040143267794f;

__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.f + erff(x * kInvSqrt2)); }
__device__ __forceinline__ float gelu_grad_f(float x) {
  return 0.5f * (1.f + erff(x * kInvSqrt2)) + x * kInvSqrt2Pi * expf(-0.5f * x * x);
}

__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  float t = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.f;
  if (w == 0) {
    t = warp_sum(t);
    if (l == 0) red[0] = t;
  }
  __syncthreads();
  return red[0];
}

// one block per row: out = LN(act(pre)) * gamma + beta ; act = GELU or identity
__global__ void __launch_bounds__(256)
act_ln_fwd_kernel(const float* __restrict__ pre, const float* __restrict__ gamma, const float* __restrict__ beta,
                  float* __restrict__ out_f32, bf16* __restrict__ out_bf16, float* __restrict__ mean_out,
                  float* __restrict__ rstd_out, int N, float eps, int apply_gelu) {
  __shared__ float red[32];
  const long row = blockIdx.x;
  const float* p = pre + row * N;
  float s = 0.f;
  for (int c = threadIdx.x; c < N; c += blockDim.x) {
    const float g = apply_gelu ? gelu_f(p[c]) : p[c];
    s += g;
  }
  const float mean = block_sum(s, red) / (float)N;
  float q = 0.f;
  for (int c = threadIdx.x; c < N; c += blockDim.x) {
    const float g = (apply_gelu ? gelu_f(p[c]) : p[c]) - mean;
    q += g * g;
  }
  const float var = block_sum(q, red) / (float)N;
  const float rstd = rsqrtf(var + eps);
  if (threadIdx.x == 0) {
    mean_out[row] = mean;
    rstd_out[row] = rstd;
  }
  for (int c = threadIdx.x; c < N; c += blockDim.x) {
    const float g = apply_gelu ? gelu_f(p[c]) : p[c];
    const float o = (g - mean) * rstd * gamma[c] + beta[c];
    if (out_f32) out_f32[row * N + c] = o;
    if (out_bf16) out_bf16[row * N + c] = __float2bfloat16_rn(o);
  }
}

// grid-stride over rows; per-thread dgamma/dbeta partials for columns tid + k*256 (N <= 256*kMaxCols)
constexpr int kMaxCols = 16;
__global__ void __launch_bounds__(256)
act_ln_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ pre, const float* __restrict__ gamma,
                  const float* __restrict__ mean_in, const float* __restrict__ rstd_in, float* __restrict__ dpre,
                  float* __restrict__ dgamma, float* __restrict__ dbeta, long M, int N, int apply_gelu) {
  __shared__ float red[32];
  float dg_acc[kMaxCols], db_acc[kMaxCols];
#pragma unroll
  for (int k = 0; k < kMaxCols; ++k) dg_acc[k] = db_acc[k] = 0.f;
  for (long row = blockIdx.x; row < M; row += gridDim.x) {
    const float* p = pre + row * N;
    const float* d = dout + row * N;
    const float mean = mean_in[row], rstd = rstd_in[row];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < kMaxCols; ++k) {
      const int c = threadIdx.x + k * 256;
      if (c < N) {
        const float g = apply_gelu ? gelu_f(p[c]) : p[c];
        const float xh = (g - mean) * rstd;
        const float dxh = d[c] * gamma[c];
        s1 += dxh;
        s2 += dxh * xh;
        dg_acc[k] += d[c] * xh;
        db_acc[k] += d[c];
      }
    }
    const float m1 = block_sum(s1, red) / (float)N;
    const float m2 = block_sum(s2, red) / (float)N;
#pragma unroll
    for (int k = 0; k < kMaxCols; ++k) {
      const int c = threadIdx.x + k * 256;
      if (c < N) {
        const float x = p[c];
        const float g = apply_gelu ? gelu_f(x) : x;
        const float xh = (g - mean) * rstd;
        const float dxh = d[c] * gamma[c];
        float dgv = rstd * (dxh - m1 - xh * m2);
        if (apply_gelu) dgv *= gelu_grad_f(x);
        dpre[row * N + c] = dgv;
      }
    }
  }
#pragma unroll
  for (int k = 0; k < kMaxCols; ++k) {
    const int c = threadIdx.x + k * 256;
    if (c < N) {
      atomicAdd(dgamma + c, dg_acc[k]);
      atomicAdd(dbeta + c, db_acc[k]);
    }
  }
}

// Wide rows (N > 256*kMaxCols, e.g. bn0 = LayerNorm(T * dirs * H) with rnn_out='all': 60 x 2 x 64 = 7680 columns): the
// same row algebra with a column loop; dgamma / dbeta leave as one atomic per (row, column) -- rows = clips, a few hundred
__global__ void __launch_bounds__(256)
act_ln_bwd_wide_kernel(const float* __restrict__ dout, const float* __restrict__ pre, const float* __restrict__ gamma,
                       const float* __restrict__ mean_in, const float* __restrict__ rstd_in, float* __restrict__ dpre,
                       float* __restrict__ dgamma, float* __restrict__ dbeta, long M, int N, int apply_gelu) {
  __shared__ float red[32];
  for (long row = blockIdx.x; row < M; row += gridDim.x) {
    const float* p = pre + row * N;
    const float* d = dout + row * N;
    const float mean = mean_in[row], rstd = rstd_in[row];
    float s1 = 0.f, s2 = 0.f;
    for (int c = threadIdx.x; c < N; c += 256) {
      const float g = apply_gelu ? gelu_f(p[c]) : p[c];
      const float xh = (g - mean) * rstd;
      const float dxh = d[c] * gamma[c];
      s1 += dxh;
      s2 += dxh * xh;
      atomicAdd(dgamma + c, d[c] * xh);
      atomicAdd(dbeta + c, d[c]);
    }
    const float m1 = block_sum(s1, red) / (float)N;
    const float m2 = block_sum(s2, red) / (float)N;
    for (int c = threadIdx.x; c < N; c += 256) {
      const float x = p[c];
      const float g = apply_gelu ? gelu_f(x) : x;
      const float xh = (g - mean) * rstd;
      const float dxh = d[c] * gamma[c];
      float dgv = rstd * (dxh - m1 - xh * m2);
      if (apply_gelu) dgv *= gelu_grad_f(x);
      dpre[row * N + c] = dgv;
    }
  }
}

// ---- fp32 SIMT GEMM: C[M,N] = alpha * op(A)[M,K] op(B)[K,N] + beta * C, row-major ----
constexpr int TS = 64, TK = 32;
template <bool TA, bool TB>
__global__ void __launch_bounds__(256)
sgemm_kernel(int M, int N, int K, float alpha, const float* __restrict__ A, long lda, const float* __restrict__ B,
             long ldb, float beta, float* __restrict__ C, long ldc, const float* __restrict__ bias,
             const float* __restrict__ bias2, int k_per_split) {
  // split-K (gridDim.z > 1): each z-slice reduces its own K range and adds into a pre-zeroed C with atomics --
  // skinny products (LSTM weight gradients: 4H x In outputs over K = B*T rows) otherwise run on 1-2 CTAs
  __shared__ float As[TK][TS + 4];
  __shared__ float Bs[TK][TS + 4];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * TS, n0 = blockIdx.x * TS;
  const int k_begin = blockIdx.z * k_per_split;
  const int k_end = min(K, k_begin + k_per_split);
  const bool split = gridDim.z > 1;
  float acc[4][4] = {};
  // software pipeline: the global loads of k-tile i+1 are in flight (registers) while tile i is multiplied --
  // the skinny LRCN products (64 x 128 outputs, K = 256..1024) run on 1-4 CTAs and were pure load latency
  constexpr int kPer = TS * TK / 256;       // elements of each operand tile per thread
  float ra[kPer], rb[kPer];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int q = 0; q < kPer; ++q) {
      const int i = threadIdx.x + q * 256;
      int m, k;
      if (TA) { m = i % TS; k = i / TS; } else { k = i % TK; m = i / TK; }
      const int gm = m0 + m, gk = k0 + k;
      ra[q] = (gm < M && gk < k_end) ? (TA ? A[(long)gk * lda + gm] : A[(long)gm * lda + gk]) : 0.f;
      int n, kk;
      if (TB) { kk = i % TK; n = i / TK; } else { n = i % TS; kk = i / TS; }
      const int gn = n0 + n, gk2 = k0 + kk;
      rb[q] = (gn < N && gk2 < k_end) ? (TB ? B[(long)gn * ldb + gk2] : B[(long)gk2 * ldb + gn]) : 0.f;
    }
  };
  fetch(k_begin);
  for (int k0 = k_begin; k0 < k_end; k0 += TK) {
#pragma unroll
    for (int q = 0; q < kPer; ++q) {
      const int i = threadIdx.x + q * 256;
      if (TA) As[i / TS][i % TS] = ra[q]; else As[i % TK][i / TK] = ra[q];
      if (TB) Bs[i % TK][i / TK] = rb[q]; else Bs[i / TS][i % TS] = rb[q];
    }
    __syncthreads();
    if (k0 + TK < k_end) fetch(k0 + TK);
#pragma unroll
    for (int k = 0; k < TK; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn >= N) continue;
      float* c = C + (long)gm * ldc + gn;
      float v = alpha * acc[i][j];
      if (blockIdx.z == 0) {
        if (bias) v += bias[gn];
        if (bias2) v += bias2[gn];
      }
      if (split) atomicAdd(c, v);
      else *c = (beta == 0.f) ? v : v + beta * (*c);
    }
  }
}

// out[c] (+)= sum_r X[r, c] ; gridDim.y row slices (> 1: atomics into a pre-zeroed / accumulated out)
__global__ void __launch_bounds__(256)
colsum_kernel(const float* __restrict__ X, long ld, long M, int N, float* __restrict__ out, int accumulate) {
  __shared__ float red[8][33];
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  const int ry = threadIdx.x >> 5;
  float s = 0.f;
  if (c < N)
    for (long r = (long)blockIdx.y * 8 + ry; r < M; r += 8L * gridDim.y) s += X[r * ld + c];
  red[ry][threadIdx.x & 31] = s;
  __syncthreads();
  if (ry == 0 && c < N) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x];
    if (gridDim.y > 1) atomicAdd(out + c, t);
    else out[c] = accumulate ? out[c] + t : t;
  }
}

__global__ void __launch_bounds__(256)
cast_f32_bf16_kernel(const float* __restrict__ src, bf16* __restrict__ dst, long n) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
    dst[i] = __float2bfloat16_rn(src[i]);
}

// dst[c, r] (bf16, ld_dst) = src[r, c] (fp32, ld_src) ; 32x32 smem tile transpose
__global__ void __launch_bounds__(256)
transpose_cast_kernel(const float* __restrict__ src, long ld_src, bf16* __restrict__ dst, long ld_dst, long R,
                      long Ccols) {
  __shared__ float tile[32][33];
  const long r0 = (long)blockIdx.y * 32, c0 = (long)blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8) {
    const long r = r0 + i, c = c0 + tx;
    tile[i][tx] = (r < R && c < Ccols) ? src[r * ld_src + c] : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const long c = c0 + i, r = r0 + tx;
    if (c < Ccols && r < R) dst[c * ld_dst + r] = __float2bfloat16_rn(tile[tx][i]);
  }
}

// stand-alone activations of the string-programmed Adapt stacks (medsos_lrcn/src/models_bidir.py:119-155: 's' nn.SiLU,
// 'g' nn.GELU (erf form), 'r' nn.ReLU) and of that file's head (F.silu after the LayerNorm): kind 0 relu, 1 gelu, 2 silu
__device__ __forceinline__ float act_f(float x, int kind) {
  if (kind == 0) return fmaxf(x, 0.f);
  if (kind == 1) return gelu_f(x);
  return x / (1.f + __expf(-x));
}
__device__ __forceinline__ float act_grad_f(float x, int kind) {
  if (kind == 0) return x > 0.f ? 1.f : 0.f;
  if (kind == 1) return gelu_grad_f(x);
  const float sg = 1.f / (1.f + __expf(-x));
  return sg * (1.f + x * (1.f - sg));
}
__global__ void __launch_bounds__(256)
act_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, long n, int kind) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) y[i] = act_f(x[i], kind);
}
__global__ void __launch_bounds__(256)
act_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, float* __restrict__ dx, long n, int kind) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
    dx[i] = dy[i] * act_grad_f(x[i], kind);
}

// dst[c, r] (bf16, ld_dst) = src[r, c] (bf16, ld_src) ; 32x32 smem tile transpose (no fp32 round trip of a bf16 operand)
__global__ void __launch_bounds__(256)
transpose_bf16_kernel(const bf16* __restrict__ src, long ld_src, bf16* __restrict__ dst, long ld_dst, long R, long Ccols) {
  __shared__ bf16 tile[32][34];
  const long r0 = (long)blockIdx.y * 32, c0 = (long)blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8) {
    const long r = r0 + i, c = c0 + tx;
    tile[i][tx] = (r < R && c < Ccols) ? src[r * ld_src + c] : __float2bfloat16_rn(0.f);
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const long c = c0 + i, r = r0 + tx;
    if (c < Ccols && r < R) dst[c * ld_dst + r] = tile[tx][i];
  }
}

// y = x * keep(i) / (1-p), keep(i) = [hash(seed, i) >= p] -- the same call with the same seed
// replays the mask for the backward pass (dx = dy * keep / (1-p)).
__device__ __forceinline__ uint32_t mix32(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return (uint32_t)((z ^ (z >> 31)) >> 32);
}
__global__ void __launch_bounds__(256)
dropout_kernel(const float* __restrict__ x, float* __restrict__ y, long n, float p, unsigned long long seed,
               const unsigned long long* __restrict__ seed_offset) {
  if (seed_offset != nullptr) seed += *seed_offset;      // device-side step counter: a captured launch draws a new mask per replay
  const float inv_keep = 1.f / (1.f - p);
  const uint32_t thresh = (uint32_t)((double)p * 4294967296.0);
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const uint32_t r = mix32(seed * 0xD1342543DE82EF95ull + (uint64_t)i);
    y[i] = (r >= thresh) ? x[i] * inv_keep : 0.f;
  }
}

}  // namespace

B2_API int b2_act_fwd_f32(const float* x, float* y, long n, int kind, void* stream) {
  B2_ARG_CHECK(x && y && n > 0 && kind >= 0 && kind <= 2, "b2_act_fwd_f32: null pointer, empty, or kind not in {0 relu, 1 gelu, 2 silu}");
  long blocks = (n + 255) / 256;
  const long cap = (long)b2_num_sms() * 8;
  act_fwd_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, (cudaStream_t)stream>>>(x, y, n, kind);
  B2_LAUNCH_CHECK("act_fwd_kernel");
  return 0;
}

B2_API int b2_act_bwd_f32(const float* dy, const float* x, float* dx, long n, int kind, void* stream) {
  B2_ARG_CHECK(dy && x && dx && n > 0 && kind >= 0 && kind <= 2, "b2_act_bwd_f32: null pointer, empty, or bad kind");
  long blocks = (n + 255) / 256;
  const long cap = (long)b2_num_sms() * 8;
  act_bwd_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, (cudaStream_t)stream>>>(dy, x, dx, n, kind);
  B2_LAUNCH_CHECK("act_bwd_kernel");
  return 0;
}

B2_API int b2_transpose_bf16(const void* src, long ld_src, void* dst, long ld_dst, long R, long C, void* stream) {
  B2_ARG_CHECK(src && dst && R > 0 && C > 0 && ld_src >= C && ld_dst >= R, "b2_transpose_bf16: bad arguments");
  dim3 grid((unsigned)((C + 31) / 32), (unsigned)((R + 31) / 32));
  B2_ARG_CHECK(grid.y <= 65535, "b2_transpose_bf16: too many rows");
  transpose_bf16_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const bf16*)src, ld_src, (bf16*)dst, ld_dst, R, C);
  B2_LAUNCH_CHECK("transpose_bf16_kernel");
  return 0;
}

B2_API int b2_dropout_f32(const float* x, float* y, long n, float p, unsigned long long seed,
                          const unsigned long long* seed_offset, void* stream) {
  B2_ARG_CHECK(x && y && n > 0, "b2_dropout_f32: null pointer or empty");
  B2_ARG_CHECK(p >= 0.f && p < 1.f, "b2_dropout_f32: p must be in [0,1)");
  long blocks = (n + 255) / 256;
  const long cap = (long)b2_num_sms() * 8;
  dropout_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, (cudaStream_t)stream>>>(x, y, n, p, seed, seed_offset);
  B2_LAUNCH_CHECK("dropout_kernel");
  return 0;
}

B2_API int b2_act_ln_fwd(const float* pre, const float* gamma, const float* beta, float* out_f32, void* out_bf16,
                         float* mean, float* rstd, long M, int N, float eps, int apply_gelu, void* stream) {
  B2_ARG_CHECK(pre && gamma && beta && mean && rstd && (out_f32 || out_bf16) && M > 0 && N > 0,
               "b2_act_ln_fwd: null pointer or empty");
  act_ln_fwd_kernel<<<(unsigned)M, 256, 0, (cudaStream_t)stream>>>(pre, gamma, beta, out_f32, (bf16*)out_bf16, mean,
                                                                   rstd, N, eps, apply_gelu);
  B2_LAUNCH_CHECK("act_ln_fwd_kernel");
  return 0;
}

// dgamma / dbeta are ACCUMULATED into (caller zeroes them)
B2_API int b2_act_ln_bwd(const float* dout, const float* pre, const float* gamma, const float* mean,
                         const float* rstd, float* dpre, float* dgamma, float* dbeta, long M, int N, int apply_gelu,
                         void* stream) {
  B2_ARG_CHECK(dout && pre && gamma && mean && rstd && dpre && dgamma && dbeta && M > 0 && N > 0,
               "b2_act_ln_bwd: null pointer or empty");
  const long cap = (long)b2_num_sms() * 2;
  if (N <= 256 * kMaxCols)
    act_ln_bwd_kernel<<<(unsigned)(M < cap ? M : cap), 256, 0, (cudaStream_t)stream>>>(dout, pre, gamma, mean, rstd,
                                                                                      dpre, dgamma, dbeta, M, N,
                                                                                      apply_gelu);
  else      // no column limit (the forward kernel has none either)
    act_ln_bwd_wide_kernel<<<(unsigned)(M < cap * 4 ? M : cap * 4), 256, 0, (cudaStream_t)stream>>>(
        dout, pre, gamma, mean, rstd, dpre, dgamma, dbeta, M, N, apply_gelu);
  B2_LAUNCH_CHECK("act_ln_bwd_kernel");
  return 0;
}

B2_API int b2_sgemm(int trans_a, int trans_b, int M, int N, int K, float alpha, const float* A, long lda,
                    const float* B, long ldb, float beta, float* C, long ldc, const float* bias, const float* bias2,
                    void* stream) {
  B2_ARG_CHECK(A && B && C && M > 0 && N > 0 && K > 0, "b2_sgemm: null pointer or empty shape");
  cudaStream_t st = (cudaStream_t)stream;
  const int tiles = b2_ceil_div(N, TS) * b2_ceil_div(M, TS);
  int splits = 1;
  // skinny output, long reduction: split K.  Only the A^T B form (weight gradients) -- atomics make the sum order
  // vary from run to run, and forward passes must stay bit-reproducible (torch.save / torch.load round trips)
  if (trans_a && beta == 0.f && tiles * 2 <= b2_num_sms() && K >= 8 * TK) {
    splits = b2_num_sms() / tiles;
    const int max_splits = K / (4 * TK);
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
  }
  int k_per_split = (b2_ceil_div(K, splits) + TK - 1) / TK * TK;
  splits = b2_ceil_div(K, k_per_split);
  if (splits > 1) B2_CUDA_CHECK(cudaMemset2DAsync(C, (size_t)ldc * 4, 0, (size_t)N * 4, (size_t)M, st));
  dim3 grid(b2_ceil_div(N, TS), b2_ceil_div(M, TS), splits);
  if (!trans_a && !trans_b) sgemm_kernel<false, false><<<grid, 256, 0, st>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, bias, bias2, k_per_split);
  else if (!trans_a && trans_b) sgemm_kernel<false, true><<<grid, 256, 0, st>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, bias, bias2, k_per_split);
  else if (trans_a && !trans_b) sgemm_kernel<true, false><<<grid, 256, 0, st>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, bias, bias2, k_per_split);
  else sgemm_kernel<true, true><<<grid, 256, 0, st>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, bias, bias2, k_per_split);
  B2_LAUNCH_CHECK("sgemm_kernel");
  return 0;
}

B2_API int b2_colsum_f32(const float* X, long ld, long M, int N, float* out, int accumulate, void* stream) {
  B2_ARG_CHECK(X && out && M > 0 && N > 0, "b2_colsum_f32: null pointer or empty");
  cudaStream_t st = (cudaStream_t)stream;
  const int col_blocks = b2_ceil_div(N, 32);
  int slices = 1;
  if (M >= 256 && col_blocks * 2 <= b2_num_sms()) {
    slices = b2_num_sms() / col_blocks;
    if (slices > M / 64) slices = (int)(M / 64);
    if (slices < 1) slices = 1;
  }
  if (slices > 1 && !accumulate) B2_CUDA_CHECK(cudaMemsetAsync(out, 0, (size_t)N * 4, st));
  colsum_kernel<<<dim3(col_blocks, slices), 256, 0, st>>>(X, ld, M, N, out, accumulate);
  B2_LAUNCH_CHECK("colsum_kernel");
  return 0;
}

B2_API int b2_cast_f32_bf16(const float* src, void* dst, long n, void* stream) {
  B2_ARG_CHECK(src && dst && n > 0, "b2_cast_f32_bf16: null pointer or empty");
  long blocks = (n + 255) / 256;
  const long cap = (long)b2_num_sms() * 8;
  cast_f32_bf16_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, (cudaStream_t)stream>>>(src, (bf16*)dst, n);
  B2_LAUNCH_CHECK("cast_f32_bf16_kernel");
  return 0;
}

B2_API int b2_transpose_cast_f32_bf16(const float* src, long ld_src, void* dst, long ld_dst, long R, long Ccols,
                                      void* stream) {
  B2_ARG_CHECK(src && dst && R > 0 && Ccols > 0, "b2_transpose_cast_f32_bf16: null pointer or empty");
  dim3 grid(b2_ceil_div(Ccols, 32), b2_ceil_div(R, 32));
  transpose_cast_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, ld_src, (bf16*)dst, ld_dst, R, Ccols);
  B2_LAUNCH_CHECK("transpose_cast_kernel");
  return 0;
}
