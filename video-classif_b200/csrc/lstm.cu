// K3 / K4 -- persistent LSTM recurrence (forward) and BPTT (backward) kernels.
//
// The input-to-gate product for ALL T steps, G = X W_ih^T + b_ih + b_hh, is hoisted out of the
// recurrence and computed once by the tcgen05 GEMM (gemm_tc.cu); these kernels run the T
// dependent steps without returning to the host.  One CTA owns NB = 4 batch rows of one
// (layer, direction): every thread keeps ONE ROW of W_hh in registers for the whole sequence
// (gate-row ownership, 4H threads), h lives in shared memory, c in registers.  Per step:
//   phase A  thread j: pre[b][j] = G[b,t,j] + <W_hh[j,:], h[b,:]>          (NB dot products)
//   phase B  thread (b,k): i,f,g,o -> c,h update, h written to the [B,T,dirs*H] output
// Gate order i,f,g,o; zero initial state; `reverse` runs t = T-1..0 (nn.LSTM `_reverse` weights).
// Reference: torch.nn.LSTM as configured at nb:169, medsos_lrcn/src/models.py:156-158, lrcn/lrcn.py:236.
//
// Backward walks the steps in the opposite order, keeps dc / dh in registers and shared memory,
// writes dG[b,t,:] for the hoisted dW_ih / dX GEMMs and accumulates dW_hh rows in registers
// (flushed once per CTA with atomics).
#include "common.cuh"

namespace {

constexpr int NB = 4;

template <int HP>  // HP >= H, register array size for one W_hh row
__global__ void __launch_bounds__(4 * HP)
lstm_fwd_kernel(const float* __restrict__ G, const float* __restrict__ Whh, float* __restrict__ out, long out_ld,
                float* __restrict__ gates, float* __restrict__ cst, int B, int T, int H, int reverse) {
  extern __shared__ float sm[];
  float* h_s = sm;                  // [NB][H]
  float* pre_s = sm + NB * H;       // [NB][4H]
  const int H4 = 4 * H;
  const int j = threadIdx.x;        // gate row
  const int b0 = blockIdx.x * NB;
  const bool active = j < H4;
  float w[HP];
#pragma unroll
  for (int k = 0; k < HP; ++k) w[k] = (active && k < H) ? Whh[(long)j * H + k] : 0.f;
  // cell-update mapping
  const int ub = j / H, uk = j - ub * H;
  const bool upd = active && ub < NB && (b0 + ub) < B;
  float c = 0.f;
  for (int i = j; i < NB * H; i += blockDim.x) h_s[i] = 0.f;
  __syncthreads();
  // the hoisted gate pre-activations of step s+1 are fetched while step s computes: the only global-memory
  // latency left on the recurrence's critical path is the first step's
  float g_next[NB];
#pragma unroll
  for (int b = 0; b < NB; ++b)
    g_next[b] = (active && b0 + b < B) ? G[((long)(b0 + b) * T + (reverse ? T - 1 : 0)) * H4 + j] : 0.f;
  for (int step = 0; step < T; ++step) {
    const int t = reverse ? T - 1 - step : step;
    float g_cur[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) g_cur[b] = g_next[b];
    if (step + 1 < T) {
      const int tn = reverse ? t - 1 : t + 1;
#pragma unroll
      for (int b = 0; b < NB; ++b)
        if (active && b0 + b < B) g_next[b] = G[((long)(b0 + b) * T + tn) * H4 + j];
    }
    if (active) {
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        if (b0 + b < B) {
          float acc = g_cur[b];
          const float* hb = h_s + b * H;
#pragma unroll
          for (int k = 0; k < HP; ++k)
            if (k < H) acc = fmaf(w[k], hb[k], acc);
          pre_s[b * H4 + j] = acc;
        }
      }
    }
    __syncthreads();
    if (upd) {
      const float* p = pre_s + ub * H4;
      const float ig = sigmoidf_(p[uk]);
      const float fg = sigmoidf_(p[H + uk]);
      const float gg = tanhf_(p[2 * H + uk]);
      const float og = sigmoidf_(p[3 * H + uk]);
      c = fg * c + ig * gg;
      const float h = og * tanhf_(c);
      h_s[ub * H + uk] = h;
      const long bt = (long)(b0 + ub) * T + t;
      out[bt * out_ld + uk] = h;
      if (gates != nullptr) {
        float* gp = gates + bt * H4;
        gp[uk] = ig;
        gp[H + uk] = fg;
        gp[2 * H + uk] = gg;
        gp[3 * H + uk] = og;
        cst[bt * H + uk] = c;
      }
    }
    __syncthreads();
  }
}

template <int HP>
__global__ void __launch_bounds__(4 * HP)
lstm_bwd_kernel(const float* __restrict__ dout, long dout_ld, const float* __restrict__ out, long out_ld,
                const float* __restrict__ gates, const float* __restrict__ cst, const float* __restrict__ Whh,
                float* __restrict__ dG, long dG_ld, float* __restrict__ dWhh, int B, int T, int H, int reverse) {
  extern __shared__ float sm[];
  const int H4 = 4 * H;
  float* w_s = sm;                         // [4H][H]
  float* dg_s = w_s + (long)H4 * H;        // [NB][4H]
  float* hp_s = dg_s + NB * H4;            // [NB][H]   h_{prev} of the current step
  const int j = threadIdx.x;
  const int b0 = blockIdx.x * NB;
  const bool active = j < H4;
  for (int i = j; i < H4 * H; i += blockDim.x) w_s[i] = Whh[i];
  const int ub = j / H, uk = j - ub * H;
  const bool upd = active && ub < NB && (b0 + ub) < B;
  float dc = 0.f, dh_rec = 0.f;   // recurrent gradients carried to the previous step (thread-private)
  float dw[HP];
#pragma unroll
  for (int k = 0; k < HP; ++k) dw[k] = 0.f;
  __syncthreads();
  // saved activations and the incoming gradient of step s-1 are fetched while step s computes
  struct Saved { float ig, fg, gg, og, cc, cprev, dout, hprev; };
  auto fetch = [&](int step) {
    Saved v = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (upd) {
      const int t = reverse ? T - 1 - step : step;
      const int tp = reverse ? t + 1 : t - 1;
      const long bt = (long)(b0 + ub) * T + t;
      const float* gp = gates + bt * H4;
      v.ig = gp[uk];
      v.fg = gp[H + uk];
      v.gg = gp[2 * H + uk];
      v.og = gp[3 * H + uk];
      v.cc = cst[bt * H + uk];
      v.dout = dout[bt * dout_ld + uk];
      if (step > 0) {
        v.cprev = cst[((long)(b0 + ub) * T + tp) * H + uk];
        v.hprev = out[((long)(b0 + ub) * T + tp) * out_ld + uk];
      }
    }
    return v;
  };
  Saved nxt = fetch(T - 1);
  for (int step = T - 1; step >= 0; --step) {
    const int t = reverse ? T - 1 - step : step;
    const Saved cur = nxt;
    if (step > 0) nxt = fetch(step - 1);
    if (upd) {
      const long bt = (long)(b0 + ub) * T + t;
      const float ig = cur.ig, fg = cur.fg, gg = cur.gg, og = cur.og;
      const float cc = cur.cc;
      const float cprev = cur.cprev;
      const float dh = cur.dout + dh_rec;
      const float tc = tanhf_(cc);
      const float dct = dc + dh * og * (1.f - tc * tc);
      float* d = dg_s + ub * H4;
      const float di = dct * gg * ig * (1.f - ig);
      const float df = dct * cprev * fg * (1.f - fg);
      const float dgg = dct * ig * (1.f - gg * gg);
      const float dgo = dh * tc * og * (1.f - og);
      d[uk] = di;
      d[H + uk] = df;
      d[2 * H + uk] = dgg;
      d[3 * H + uk] = dgo;
      float* go = dG + bt * dG_ld;
      go[uk] = di;
      go[H + uk] = df;
      go[2 * H + uk] = dgg;
      go[3 * H + uk] = dgo;
      dc = dct * fg;
      hp_s[ub * H + uk] = cur.hprev;
    } else if (active && ub < NB) {
      float* d = dg_s + ub * H4;   // rows beyond the batch contribute nothing
      d[uk] = d[H + uk] = d[2 * H + uk] = d[3 * H + uk] = 0.f;
      hp_s[ub * H + uk] = 0.f;
    }
    __syncthreads();
    if (active) {
      // dW_hh[j,:] += sum_b dg[b][j] * h_prev[b][:]
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        const float g = dg_s[b * H4 + j];
        const float* hb = hp_s + b * H;
#pragma unroll
        for (int k = 0; k < HP; ++k)
          if (k < H) dw[k] = fmaf(g, hb[k], dw[k]);
      }
    }
    if (upd) {
      // dh_prev[b][k] = sum_j dg[b][j] * W_hh[j][k]
      const float* d = dg_s + ub * H4;
      float acc = 0.f;
      for (int jj = 0; jj < H4; ++jj) acc = fmaf(d[jj], w_s[jj * H + uk], acc);
      dh_rec = acc;
    }
    __syncthreads();
  }
  if (active) {
#pragma unroll
    for (int k = 0; k < HP; ++k)
      if (k < H) atomicAdd(dWhh + (long)j * H + k, dw[k]);
  }
}

}  // namespace

B2_API int b2_lstm_seq_fwd(const float* G, const float* Whh, float* out, long out_ld, float* gates, float* cstate,
                           int B, int T, int H, int reverse, void* stream) {
  B2_ARG_CHECK(G && Whh && out && B > 0 && T > 0 && H > 0, "b2_lstm_seq_fwd: null pointer or empty");
  B2_ARG_CHECK(H <= 64, "b2_lstm_seq_fwd: hidden size %d > 64 is not supported yet", H);
  B2_ARG_CHECK((gates == nullptr) == (cstate == nullptr), "b2_lstm_seq_fwd: gates and cstate go together");
  const int grid = b2_ceil_div(B, NB);
  const size_t smem = (size_t)(NB * H + NB * 4 * H) * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
  if (H <= 32)
    lstm_fwd_kernel<32><<<grid, 128, smem, st>>>(G, Whh, out, out_ld, gates, cstate, B, T, H, reverse);
  else
    lstm_fwd_kernel<64><<<grid, 256, smem, st>>>(G, Whh, out, out_ld, gates, cstate, B, T, H, reverse);
  B2_LAUNCH_CHECK("lstm_fwd_kernel");
  return 0;
}

// dWhh is ACCUMULATED into (caller zeroes it); dG is fully overwritten.
B2_API int b2_lstm_seq_bwd(const float* dout, long dout_ld, const float* out, long out_ld, const float* gates,
                           const float* cstate, const float* Whh, float* dG, long dG_ld, float* dWhh, int B, int T,
                           int H, int reverse, void* stream) {
  B2_ARG_CHECK(dout && out && gates && cstate && Whh && dG && dWhh && B > 0 && T > 0 && H > 0,
               "b2_lstm_seq_bwd: null pointer or empty");
  B2_ARG_CHECK(H <= 64, "b2_lstm_seq_bwd: hidden size %d > 64 is not supported yet", H);
  const int grid = b2_ceil_div(B, NB);
  const size_t smem = (size_t)(4 * H * H + NB * 4 * H + NB * H) * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
  if (H <= 32) {
    lstm_bwd_kernel<32><<<grid, 128, smem, st>>>(dout, dout_ld, out, out_ld, gates, cstate, Whh, dG, dG_ld, dWhh, B, T, H,
                                                 reverse);
  } else {
    static B2PerDeviceOnce attr;
    if (attr.needed()) {
      B2_CUDA_CHECK(cudaFuncSetAttribute(lstm_bwd_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
      attr.mark();
    }
    lstm_bwd_kernel<64><<<grid, 256, smem, st>>>(dout, dout_ld, out, out_ld, gates, cstate, Whh, dG, dG_ld, dWhh, B, T, H,
                                                 reverse);
  }
  B2_LAUNCH_CHECK("lstm_bwd_kernel");
  return 0;
}
