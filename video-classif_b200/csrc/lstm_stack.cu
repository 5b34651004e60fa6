// Whole-stack persistent LSTM: every layer and every timestep of a unidirectional nn.LSTM in ONE launch
// (forward) and one launch (BPTT), for the narrow stacks the LRCN heads use (medsos: 8 -> 32 x 3 layers,
// models.py:156-158 with all_config.py:14-17; H <= 64, input width <= 64, T <= 64).
//
// One CTA owns one batch row for the whole stack; thread j owns gate row j (4H threads, gate order i,f,g,o):
// its rows of W_ih and W_hh live in registers across all T steps of a layer, the layer's input and output
// sequences live in shared memory, so layer k+1 starts from on-chip data and the input-to-gate products need
// no separate GEMM launch at these widths.  The per-layer kernels (lstm.cu) + hoisted tcgen05 gate GEMM remain the
// path for wide inputs (small-CNN LRCN: 16384 -> 32) and bidirectional stacks.
//
// Backward walks the layers top-down and the steps in reverse: dc / dh carried in registers, the gate
// gradients of a step broadcast through shared memory; thread j accumulates ITS rows of dW_ih / dW_hh / db in
// registers over the whole sequence (flushed once per layer with atomics), W_ih / W_hh sit in shared memory for
// the transposed products (dh_{t-1} = dG W_hh, dx_t = dG W_ih), and dx of layer k is the incoming gradient of
// layer k-1 without leaving the SM.
#include "common.cuh"

namespace {

constexpr int kMaxLayers = 8;
constexpr int kMaxIn = 64;

struct StackParams {
  int layers, B, T, H, In0;
  const float* x;                 // [B, T, In0]
  const float* w_ih[kMaxLayers];  // [4H, In_l]
  const float* w_hh[kMaxLayers];  // [4H, H]
  const float* b_ih[kMaxLayers];  // [4H]
  const float* b_hh[kMaxLayers];
  float* out;                     // [layers, B, T, H]   every layer's output sequence
  float* gates;                   // [layers, B, T, 4H]  or null (inference)
  float* cst;                     // [layers, B, T, H]   or null
};

struct StackGradParams {
  const float* dout;              // [B, T, H] gradient of the top layer's output sequence
  float* dx;                      // [B, T, In0] or null
  float* dw_ih[kMaxLayers];       // ACCUMULATED into (caller zeroes)
  float* dw_hh[kMaxLayers];
  float* db[kMaxLayers];          // [4H]: gradient of b_ih (== gradient of b_hh)
};

// dst[r * dst_stride + k] = src[r * cols + k] for a dense [rows, cols] global matrix, by the whole block
__device__ __forceinline__ void stage_rows(float* dst, const float* __restrict__ src, int rows, int cols, int dst_stride) {
  const int total = rows * cols;
  const int nt = blockDim.x;
  for (int i0 = 0; i0 < total; i0 += 8 * nt) {
    float v[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int i = i0 + q * nt + threadIdx.x;
      v[q] = i < total ? src[i] : 0.f;
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int i = i0 + q * nt + threadIdx.x;
      if (i < total) {
        const int r = i / cols;
        dst[r * dst_stride + (i - r * cols)] = v[q];
      }
    }
  }
}

template <int HP>   // HP >= H: register array size of a W_hh row
__global__ void __launch_bounds__(4 * HP)
lstm_stack_fwd_kernel(StackParams p) {
  extern __shared__ float sm[];
  const int T = p.T, H = p.H, H4 = 4 * H;
  float* seq_a = sm;                         // [T][kMaxIn]  input sequence of the current layer
  float* seq_b = seq_a + T * kMaxIn;         // [T][kMaxIn]  output sequence of the current layer
  float* pre_s = seq_b + T * kMaxIn;         // [4H]
  float* h_s = pre_s + 4 * HP;               // [H]
  const int j = threadIdx.x;
  const int b = blockIdx.x;
  const bool active = j < H4;
  const bool upd = j < H;
  // padding columns must read as zero: the dot products below run over whole float4 groups
  for (int i = j; i < 2 * T * kMaxIn; i += blockDim.x) seq_a[i] = 0.f;
  for (int i = j; i < HP; i += blockDim.x) h_s[i] = 0.f;
  __syncthreads();
  stage_rows(seq_a, p.x + (long)b * T * p.In0, T, p.In0, kMaxIn);
  int In = p.In0;
  for (int l = 0; l < p.layers; ++l) {
    float wih[kMaxIn], whh[HP];
#pragma unroll
    for (int k = 0; k < kMaxIn; ++k) wih[k] = (active && k < In) ? p.w_ih[l][(long)j * In + k] : 0.f;
#pragma unroll
    for (int k = 0; k < HP; ++k) whh[k] = (active && k < H) ? p.w_hh[l][(long)j * H + k] : 0.f;
    const float bias = active ? p.b_ih[l][j] + p.b_hh[l][j] : 0.f;
    float c = 0.f;
    if (upd) h_s[j] = 0.f;
    __syncthreads();
    const long lb = ((long)l * p.B + b) * T;
    for (int t = 0; t < T; ++t) {
      if (active) {
        // 128-bit broadcast loads, four in flight, four independent partial sums (with scalar loads the
        // compiler recycled one destination register and every element paid a full LDS latency: 2.2 us / step)
        float acc[4] = {bias, 0.f, 0.f, 0.f};
        const float4* xt4 = reinterpret_cast<const float4*>(seq_a + t * kMaxIn);
#pragma unroll
        for (int k0 = 0; k0 < kMaxIn; k0 += 16) {
          if (k0 < In) {
            float4 v[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) v[q] = xt4[k0 / 4 + q];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              acc[0] = fmaf(wih[k0 + 4 * q], v[q].x, acc[0]);
              acc[1] = fmaf(wih[k0 + 4 * q + 1], v[q].y, acc[1]);
              acc[2] = fmaf(wih[k0 + 4 * q + 2], v[q].z, acc[2]);
              acc[3] = fmaf(wih[k0 + 4 * q + 3], v[q].w, acc[3]);
            }
          }
        }
        const float4* h4 = reinterpret_cast<const float4*>(h_s);
#pragma unroll
        for (int k0 = 0; k0 < HP; k0 += 16) {
          if (k0 < H) {
            float4 v[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) v[q] = h4[k0 / 4 + q];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              acc[0] = fmaf(whh[k0 + 4 * q], v[q].x, acc[0]);
              acc[1] = fmaf(whh[k0 + 4 * q + 1], v[q].y, acc[1]);
              acc[2] = fmaf(whh[k0 + 4 * q + 2], v[q].z, acc[2]);
              acc[3] = fmaf(whh[k0 + 4 * q + 3], v[q].w, acc[3]);
            }
          }
        }
        pre_s[j] = (acc[0] + acc[1]) + (acc[2] + acc[3]);
      }
      __syncthreads();
      if (upd) {
        const float ig = sigmoidf_(pre_s[j]);
        const float fg = sigmoidf_(pre_s[H + j]);
        const float gg = tanhf_(pre_s[2 * H + j]);
        const float og = sigmoidf_(pre_s[3 * H + j]);
        c = fg * c + ig * gg;
        const float h = og * tanhf_(c);
        h_s[j] = h;
        seq_b[t * kMaxIn + j] = h;
        p.out[(lb + t) * H + j] = h;
        if (p.gates != nullptr) {
          float* gp = p.gates + (lb + t) * H4;
          gp[j] = ig;
          gp[H + j] = fg;
          gp[2 * H + j] = gg;
          gp[3 * H + j] = og;
          p.cst[(lb + t) * H + j] = c;
        }
      }
      __syncthreads();
    }
    float* tmp = seq_a;   // this layer's output is the next layer's input
    seq_a = seq_b;
    seq_b = tmp;
    In = H;
  }
}

template <int HP>
__global__ void __launch_bounds__(4 * HP)
lstm_stack_bwd_kernel(StackParams p, StackGradParams g) {
  extern __shared__ float sm[];
  const int T = p.T, H = p.H, H4 = 4 * H;
  constexpr int kWarps = 4 * HP / 32;
  float* xin_s = sm;                          // [T][kMaxIn]  input sequence of the current layer
  float* hout_s = xin_s + T * kMaxIn;         // [T][kMaxIn]  output sequence of the current layer (h_{t-1} lookups)
  float* dh_s = hout_s + T * kMaxIn;          // [T][kMaxIn]  incoming gradient of the output sequence
  float* dxs_s = dh_s + T * kMaxIn;           // [T][kMaxIn]  gradient of the input sequence (next layer down)
  float* wih_s = dxs_s + T * kMaxIn;          // [4H][kMaxIn]
  float* whh_s = wih_s + 4 * HP * kMaxIn;     // [4H][HP]
  float* dg_s = whh_s + 4 * HP * HP;          // [4H]
  float* part_s = dg_s + 4 * HP;              // [kWarps][2 * kMaxIn]: per-warp partial dh_{t-1} | dx_t
  const int j = threadIdx.x;
  const int warp = j >> 5, lane = j & 31;
  const int b = blockIdx.x;
  const bool active = j < H4;
  const bool upd = j < H;
  // rows of the gate-gradient vector this warp reduces over in the transposed products
  const int rows_per_warp = (H4 + kWarps - 1) / kWarps;
  const int r_begin = warp * rows_per_warp;
  const int r_end = min(H4, r_begin + rows_per_warp);
  for (int i = j; i < 2 * T * kMaxIn; i += blockDim.x) xin_s[i] = 0.f;     // xin_s | hout_s: zero padding columns
  for (int i = j; i < T * H; i += blockDim.x) {
    const int t = i / H, k = i - t * H;
    dh_s[t * kMaxIn + k] = g.dout[((long)b * T + t) * H + k];
  }
  for (int l = p.layers - 1; l >= 0; --l) {
    const int In = l == 0 ? p.In0 : H;
    const long lb = ((long)l * p.B + b) * T;
    __syncthreads();     // previous layer's use of the staging buffers is over
    // staging: batches of 8 independent coalesced loads per thread (latency paid once per batch, not per element)
    stage_rows(xin_s, l == 0 ? p.x + (long)b * T * In : p.out + (((long)(l - 1)) * p.B + b) * T * H, T, In, kMaxIn);
    stage_rows(hout_s, p.out + lb * H, T, H, kMaxIn);
    stage_rows(wih_s, p.w_ih[l], H4, In, kMaxIn);
    stage_rows(whh_s, p.w_hh[l], H4, H, HP);
    float dwih[kMaxIn], dwhh[HP];
#pragma unroll
    for (int k = 0; k < kMaxIn; ++k) dwih[k] = 0.f;
#pragma unroll
    for (int k = 0; k < HP; ++k) dwhh[k] = 0.f;
    float dbias = 0.f, dc = 0.f, dh_rec = 0.f;
    // saved gates / cell states of step t-1 are fetched while step t computes
    float nig = 0.f, nfg = 0.f, ngg = 0.f, nog = 0.f, ncc = 0.f, ncp = 0.f;
    auto fetch = [&](int t) {
      const float* gp = p.gates + (lb + t) * H4;
      nig = gp[j]; nfg = gp[H + j]; ngg = gp[2 * H + j]; nog = gp[3 * H + j];
      ncc = p.cst[(lb + t) * H + j];
      ncp = t > 0 ? p.cst[(lb + t - 1) * H + j] : 0.f;
    };
    if (upd) fetch(T - 1);
    __syncthreads();
    for (int t = T - 1; t >= 0; --t) {
      if (upd) {
        const float ig = nig, fg = nfg, gg = ngg, og = nog, cc = ncc, cprev = ncp;
        if (t > 0) fetch(t - 1);
        const float dh = dh_s[t * kMaxIn + j] + dh_rec;
        const float tc = tanhf_(cc);
        const float dct = dc + dh * og * (1.f - tc * tc);
        dg_s[j] = dct * gg * ig * (1.f - ig);
        dg_s[H + j] = dct * cprev * fg * (1.f - fg);
        dg_s[2 * H + j] = dct * ig * (1.f - gg * gg);
        dg_s[3 * H + j] = dh * tc * og * (1.f - og);
        dc = dct * fg;
      }
      __syncthreads();
      if (active) {
        const float dgj = dg_s[j];
        const float4* xt4 = reinterpret_cast<const float4*>(xin_s + t * kMaxIn);
#pragma unroll
        for (int k0 = 0; k0 < kMaxIn; k0 += 16) {
          if (k0 < In) {
            float4 v[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) v[q] = xt4[k0 / 4 + q];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              dwih[k0 + 4 * q] = fmaf(dgj, v[q].x, dwih[k0 + 4 * q]);
              dwih[k0 + 4 * q + 1] = fmaf(dgj, v[q].y, dwih[k0 + 4 * q + 1]);
              dwih[k0 + 4 * q + 2] = fmaf(dgj, v[q].z, dwih[k0 + 4 * q + 2]);
              dwih[k0 + 4 * q + 3] = fmaf(dgj, v[q].w, dwih[k0 + 4 * q + 3]);
            }
          }
        }
        if (t > 0) {
          const float4* hp4 = reinterpret_cast<const float4*>(hout_s + (t - 1) * kMaxIn);
#pragma unroll
          for (int k0 = 0; k0 < HP; k0 += 16) {
            if (k0 < H) {
              float4 v[4];
#pragma unroll
              for (int q = 0; q < 4; ++q) v[q] = hp4[k0 / 4 + q];
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                dwhh[k0 + 4 * q] = fmaf(dgj, v[q].x, dwhh[k0 + 4 * q]);
                dwhh[k0 + 4 * q + 1] = fmaf(dgj, v[q].y, dwhh[k0 + 4 * q + 1]);
                dwhh[k0 + 4 * q + 2] = fmaf(dgj, v[q].z, dwhh[k0 + 4 * q + 2]);
                dwhh[k0 + 4 * q + 3] = fmaf(dgj, v[q].w, dwhh[k0 + 4 * q + 3]);
              }
            }
          }
        }
        dbias += dgj;
      }
      {
        // transposed products, split over the warps by gate row: dh_{t-1}[k] = sum_r dG[r] W_hh[r][k],
        // dx_t[k] = sum_r dG[r] W_ih[r][k]; lane handles columns lane and lane + 32
        float a0 = 0.f, a1 = 0.f, x0 = 0.f, x1 = 0.f;
#pragma unroll 8
        for (int r = r_begin; r < r_end; ++r) {
          const float d = dg_s[r];
          a0 = fmaf(d, whh_s[r * HP + lane], a0);
          if (HP > 32) a1 = fmaf(d, whh_s[r * HP + 32 + lane], a1);
          x0 = fmaf(d, wih_s[r * kMaxIn + lane], x0);
          x1 = fmaf(d, wih_s[r * kMaxIn + 32 + lane], x1);
        }
        float* ps = part_s + warp * 2 * kMaxIn;
        ps[lane] = a0;
        ps[32 + lane] = a1;
        ps[kMaxIn + lane] = x0;
        ps[kMaxIn + 32 + lane] = x1;
      }
      __syncthreads();
      if (upd) {
        float acc = 0.f;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) acc += part_s[w * 2 * kMaxIn + j];
        dh_rec = acc;
      }
      if (j < In) {
        float acc = 0.f;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) acc += part_s[w * 2 * kMaxIn + kMaxIn + j];
        dxs_s[t * kMaxIn + j] = acc;
      }
    }
    if (active) {
#pragma unroll
      for (int k = 0; k < kMaxIn; ++k)
        if (k < In) atomicAdd(g.dw_ih[l] + (long)j * In + k, dwih[k]);
#pragma unroll
      for (int k = 0; k < HP; ++k)
        if (k < H) atomicAdd(g.dw_hh[l] + (long)j * H + k, dwhh[k]);
      atomicAdd(g.db[l] + j, dbias);
    }
    __syncthreads();      // dxs_s complete
    if (l == 0) {
      if (g.dx != nullptr)
        for (int i = j; i < T * In; i += blockDim.x) {
          const int t = i / In, k = i - t * In;
          g.dx[((long)b * T + t) * In + k] = dxs_s[t * kMaxIn + k];
        }
    } else {
      float* tmp = dh_s;    // dx of this layer is the output gradient of the layer below
      dh_s = dxs_s;
      dxs_s = tmp;
    }
  }
}

int check_stack(const char* who, int layers, int B, int T, int H, int In0) {
  B2_ARG_CHECK(layers >= 1 && layers <= kMaxLayers, "%s: 1..%d layers (got %d)", who, kMaxLayers, layers);
  B2_ARG_CHECK(B > 0 && T > 0 && T <= 64, "%s: batch > 0 and 1 <= T <= 64 (got B=%d T=%d)", who, B, T);
  B2_ARG_CHECK(H >= 1 && H <= 64 && In0 >= 1 && In0 <= kMaxIn, "%s: hidden <= 64 and input width <= %d (got H=%d In=%d)",
               who, kMaxIn, H, In0);
  return 0;
}

StackParams pack(const float* x, int In0, const void* const* w_ih, const void* const* w_hh, const void* const* b_ih,
                 const void* const* b_hh, int layers, float* out, float* gates, float* cst, int B, int T, int H) {
  StackParams p = {};
  p.layers = layers; p.B = B; p.T = T; p.H = H; p.In0 = In0;
  p.x = x; p.out = out; p.gates = gates; p.cst = cst;
  for (int l = 0; l < layers; ++l) {
    p.w_ih[l] = (const float*)w_ih[l];
    p.w_hh[l] = (const float*)w_hh[l];
    p.b_ih[l] = (const float*)b_ih[l];
    p.b_hh[l] = (const float*)b_hh[l];
  }
  return p;
}

}  // namespace

// see include/b200lrcn.h ; w_ih / w_hh / b_ih / b_hh are HOST arrays of `layers` device pointers
B2_API int b2_lstm_stack_fwd(const float* x, int In0, const void* const* w_ih, const void* const* w_hh,
                             const void* const* b_ih, const void* const* b_hh, int layers, float* out, float* gates,
                             float* cstate, int B, int T, int H, void* stream) {
  const char* who = "b2_lstm_stack_fwd";
  B2_ARG_CHECK(x && w_ih && w_hh && b_ih && b_hh && out, "%s: null pointer", who);
  B2_ARG_CHECK((gates == nullptr) == (cstate == nullptr), "%s: gates and cstate go together", who);
  if (int r = check_stack(who, layers, B, T, H, In0)) return r;
  for (int l = 0; l < layers; ++l)
    B2_ARG_CHECK(w_ih[l] && w_hh[l] && b_ih[l] && b_hh[l], "%s: null parameter pointer in layer %d", who, l);
  const StackParams p = pack(x, In0, w_ih, w_hh, b_ih, b_hh, layers, out, gates, cstate, B, T, H);
  cudaStream_t st = (cudaStream_t)stream;
  if (H <= 32) {
    const size_t smem = (size_t)(2 * T * kMaxIn + 4 * 32 + 32) * sizeof(float);
    lstm_stack_fwd_kernel<32><<<B, 128, smem, st>>>(p);
  } else {
    const size_t smem = (size_t)(2 * T * kMaxIn + 4 * 64 + 64) * sizeof(float);
    lstm_stack_fwd_kernel<64><<<B, 256, smem, st>>>(p);
  }
  B2_LAUNCH_CHECK("lstm_stack_fwd_kernel");
  return 0;
}

// dw_ih / dw_hh / db: HOST arrays of `layers` device pointers, ACCUMULATED into (caller zeroes them); dx may be NULL
B2_API int b2_lstm_stack_bwd(const float* dout, const float* x, int In0, const void* const* w_ih,
                             const void* const* w_hh, int layers, const float* out, const float* gates,
                             const float* cstate, float* dx, void* const* dw_ih, void* const* dw_hh, void* const* db,
                             int B, int T, int H, void* stream) {
  const char* who = "b2_lstm_stack_bwd";
  B2_ARG_CHECK(dout && x && w_ih && w_hh && out && gates && cstate && dw_ih && dw_hh && db, "%s: null pointer", who);
  if (int r = check_stack(who, layers, B, T, H, In0)) return r;
  StackParams p = pack(x, In0, w_ih, w_hh, w_hh, w_hh, layers, const_cast<float*>(out), const_cast<float*>(gates),
                       const_cast<float*>(cstate), B, T, H);   // biases are not read by the backward pass
  StackGradParams g = {};
  g.dout = dout;
  g.dx = dx;
  for (int l = 0; l < layers; ++l) {
    B2_ARG_CHECK(w_ih[l] && w_hh[l] && dw_ih[l] && dw_hh[l] && db[l], "%s: null pointer in layer %d", who, l);
    g.dw_ih[l] = (float*)dw_ih[l];
    g.dw_hh[l] = (float*)dw_hh[l];
    g.db[l] = (float*)db[l];
  }
  cudaStream_t st = (cudaStream_t)stream;
  static B2PerDeviceOnce attr;
  if (attr.needed()) {
    B2_CUDA_CHECK(cudaFuncSetAttribute(lstm_stack_bwd_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    B2_CUDA_CHECK(cudaFuncSetAttribute(lstm_stack_bwd_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    attr.mark();
  }
  if (H <= 32) {
    const size_t smem = (size_t)(4 * T * kMaxIn + 4 * 32 * kMaxIn + 4 * 32 * 32 + 4 * 32 + 4 * 2 * kMaxIn) * sizeof(float);
    lstm_stack_bwd_kernel<32><<<B, 128, smem, st>>>(p, g);
  } else {
    const size_t smem = (size_t)(4 * T * kMaxIn + 4 * 64 * kMaxIn + 4 * 64 * 64 + 4 * 64 + 8 * 2 * kMaxIn) * sizeof(float);
    B2_ARG_CHECK(smem <= 220 * 1024, "%s: T=%d too long for the shared-memory staging at H=%d", who, T, H);
    lstm_stack_bwd_kernel<64><<<B, 256, smem, st>>>(p, g);
  }
  B2_LAUNCH_CHECK("lstm_stack_bwd_kernel");
  return 0;
}
