// Persistent GRU recurrence (forward) and BPTT (backward) -- the temporal layer of the reference's GRU variants
// (lrcn/backup_ucf50.py:126 `LRCN2`: bidirectional nn.GRU over the small-CNN features; medsos models.py:160-170
// rnn_type="gru").  torch.nn.GRU semantics, gate order r, z, n along dim 0 of weight_ih / weight_hh [3H, .]:
//     r = sig(gi_r + hh_r)   z = sig(gi_z + hh_z)   n = tanh(gi_n + r * hh_n)   h' = (1 - z) * n + z * h
//     gi = x W_ih^T + b_ih   (hoisted for all T steps onto the tcgen05 GEMM, like the LSTM gate GEMM)
//     hh = h W_hh^T + b_hh   (b_hn sits INSIDE the r product, so b_hh cannot be folded into the hoisted GEMM)
// One CTA (4 HP threads) owns NB = 4 batch rows of one (layer, direction); thread j < 3H owns row j of W_hh (registers) for the whole
// sequence; h in shared memory.  The next step's gi is prefetched while the current step computes; the recurrent
// dot products use 128-bit broadcast loads, four in flight (see lstm_stack.cu for why).
// Backward keeps dh in registers, writes dgi for the hoisted dW_ih / dX GEMMs, accumulates dW_hh rows and db_hh in
// registers (one atomic flush per CTA).
#include "common.cuh"

namespace {

constexpr int NB = 4;

template <int HP>   // HP >= H (multiple of 16): register array size of one W_hh row
__global__ void __launch_bounds__(4 * HP)
gru_fwd_kernel(const float* __restrict__ G, const float* __restrict__ Whh, const float* __restrict__ bhh,
               float* __restrict__ out, long out_ld, float* __restrict__ saved, int B, int T, int H, int reverse) {
  extern __shared__ float sm[];
  const int H3 = 3 * H;
  float* h_s = sm;                    // [NB][HP]
  float* gi_s = h_s + NB * HP;        // [NB][3H]
  float* hh_s = gi_s + NB * H3;       // [NB][3H]
  const int j = threadIdx.x;
  const int b0 = blockIdx.x * NB;
  const bool active = j < H3;
  float w[HP];
#pragma unroll
  for (int k = 0; k < HP; ++k) w[k] = (active && k < H) ? Whh[(long)j * H + k] : 0.f;
  const float bj = active ? bhh[j] : 0.f;
  for (int i = j; i < NB * HP; i += blockDim.x) h_s[i] = 0.f;
  const int ub = j / H, uk = j - ub * H;            // update mapping: thread -> (batch row, hidden unit)
  const bool upd = ub < NB && (b0 + ub) < B && j < NB * H;
  float g_next[NB];
#pragma unroll
  for (int b = 0; b < NB; ++b)
    g_next[b] = (active && b0 + b < B) ? G[((long)(b0 + b) * T + (reverse ? T - 1 : 0)) * H3 + j] : 0.f;
  __syncthreads();
  for (int step = 0; step < T; ++step) {
    const int t = reverse ? T - 1 - step : step;
    float g_cur[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) g_cur[b] = g_next[b];
    if (step + 1 < T) {
      const int tn = reverse ? t - 1 : t + 1;
#pragma unroll
      for (int b = 0; b < NB; ++b)
        if (active && b0 + b < B) g_next[b] = G[((long)(b0 + b) * T + tn) * H3 + j];
    }
    if (active) {
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        float acc[4] = {bj, 0.f, 0.f, 0.f};
        const float4* h4 = reinterpret_cast<const float4*>(h_s + b * HP);
#pragma unroll
        for (int k0 = 0; k0 < HP; k0 += 16) {
          if (k0 < H) {
            float4 v[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) v[q] = h4[k0 / 4 + q];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              acc[0] = fmaf(w[k0 + 4 * q], v[q].x, acc[0]);
              acc[1] = fmaf(w[k0 + 4 * q + 1], v[q].y, acc[1]);
              acc[2] = fmaf(w[k0 + 4 * q + 2], v[q].z, acc[2]);
              acc[3] = fmaf(w[k0 + 4 * q + 3], v[q].w, acc[3]);
            }
          }
        }
        hh_s[b * H3 + j] = (acc[0] + acc[1]) + (acc[2] + acc[3]);
        gi_s[b * H3 + j] = g_cur[b];
      }
    }
    __syncthreads();
    if (upd) {
      const float* gi = gi_s + ub * H3;
      const float* hh = hh_s + ub * H3;
      const float r = sigmoidf_(gi[uk] + hh[uk]);
      const float z = sigmoidf_(gi[H + uk] + hh[H + uk]);
      const float hn = hh[2 * H + uk];
      const float n = tanhf_(gi[2 * H + uk] + r * hn);
      const float hp = h_s[ub * HP + uk];
      const float h = (1.f - z) * n + z * hp;
      const long bt = (long)(b0 + ub) * T + t;
      out[bt * out_ld + uk] = h;
      if (saved != nullptr) {
        float* sp = saved + bt * 4 * H;
        sp[uk] = r;
        sp[H + uk] = z;
        sp[2 * H + uk] = n;
        sp[3 * H + uk] = hn;
      }
      h_s[ub * HP + uk] = h;       // own slot only: the other threads read h_s in the NEXT step's dot products
    }
    __syncthreads();
  }
}

template <int HP>
__global__ void __launch_bounds__(4 * HP)
gru_bwd_kernel(const float* __restrict__ dout, long dout_ld, const float* __restrict__ out, long out_ld,
               const float* __restrict__ saved, const float* __restrict__ Whh, float* __restrict__ dG, long dG_ld,
               float* __restrict__ dWhh, float* __restrict__ dbhh, int B, int T, int H, int reverse) {
  extern __shared__ float sm[];
  const int H3 = 3 * H;
  float* w_s = sm;                          // [3H][HP]
  float* dhh_s = w_s + 3 * HP * HP;         // [NB][3 HP]   gradient of hh = W_hh h + b_hh
  float* hp_s = dhh_s + NB * 3 * HP;        // [NB][HP]     h_{t-1}
  const int j = threadIdx.x;
  const int b0 = blockIdx.x * NB;
  const bool active = j < H3;
  for (int i = j; i < 3 * HP * HP; i += blockDim.x) {
    const int r = i / HP, k = i - r * HP;
    w_s[i] = (r < H3 && k < H) ? Whh[(long)r * H + k] : 0.f;
  }
  for (int i = j; i < NB * HP; i += blockDim.x) hp_s[i] = 0.f;
  const int ub = j / H, uk = j - ub * H;
  const bool urow = ub < NB && j < NB * H;
  const bool upd = urow && (b0 + ub) < B;
  float dh_rec = 0.f;
  float dw[HP];
#pragma unroll
  for (int k = 0; k < HP; ++k) dw[k] = 0.f;
  float db = 0.f;
  __syncthreads();
  for (int step = T - 1; step >= 0; --step) {
    const int t = reverse ? T - 1 - step : step;
    const int tp = reverse ? t + 1 : t - 1;       // time index of h_{prev} in recurrence order
    if (upd) {
      const long bt = (long)(b0 + ub) * T + t;
      const float* sp = saved + bt * 4 * H;
      const float r = sp[uk], z = sp[H + uk], n = sp[2 * H + uk], hn = sp[3 * H + uk];
      const float hprev = step > 0 ? out[((long)(b0 + ub) * T + tp) * out_ld + uk] : 0.f;
      const float dh = dout[bt * dout_ld + uk] + dh_rec;
      const float dn_pre = dh * (1.f - z) * (1.f - n * n);
      const float dz_pre = dh * (hprev - n) * z * (1.f - z);
      const float dr_pre = dn_pre * hn * r * (1.f - r);
      float* go = dG + bt * dG_ld;
      go[uk] = dr_pre;
      go[H + uk] = dz_pre;
      go[2 * H + uk] = dn_pre;
      float* d = dhh_s + ub * 3 * HP;
      d[uk] = dr_pre;
      d[H + uk] = dz_pre;
      d[2 * H + uk] = dn_pre * r;
      hp_s[ub * HP + uk] = hprev;
      dh_rec = dh * z;                 // direct path; the W_hh path is added after the barrier
    } else if (urow) {
      float* d = dhh_s + ub * 3 * HP;  // rows beyond the batch contribute nothing
      d[uk] = d[H + uk] = d[2 * H + uk] = 0.f;
      hp_s[ub * HP + uk] = 0.f;
    }
    __syncthreads();
    if (active) {
      // dW_hh[j,:] += sum_b dhh[b][j] * h_prev[b][:] ; db_hh[j] += sum_b dhh[b][j]
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        const float gj = dhh_s[b * 3 * HP + j];
        db += gj;
        const float4* h4 = reinterpret_cast<const float4*>(hp_s + b * HP);
#pragma unroll
        for (int k0 = 0; k0 < HP; k0 += 16) {
          if (k0 < H) {
            float4 v[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) v[q] = h4[k0 / 4 + q];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              dw[k0 + 4 * q] = fmaf(gj, v[q].x, dw[k0 + 4 * q]);
              dw[k0 + 4 * q + 1] = fmaf(gj, v[q].y, dw[k0 + 4 * q + 1]);
              dw[k0 + 4 * q + 2] = fmaf(gj, v[q].z, dw[k0 + 4 * q + 2]);
              dw[k0 + 4 * q + 3] = fmaf(gj, v[q].w, dw[k0 + 4 * q + 3]);
            }
          }
        }
      }
    }
    if (upd) {
      // dh_prev[b][k] += sum_r dhh[b][r] * W_hh[r][k]
      const float* d = dhh_s + ub * 3 * HP;
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      int r = 0;
      for (; r + 4 <= H3; r += 4) {
        a0 = fmaf(d[r], w_s[r * HP + uk], a0);
        a1 = fmaf(d[r + 1], w_s[(r + 1) * HP + uk], a1);
        a2 = fmaf(d[r + 2], w_s[(r + 2) * HP + uk], a2);
        a3 = fmaf(d[r + 3], w_s[(r + 3) * HP + uk], a3);
      }
      for (; r < H3; ++r) a0 = fmaf(d[r], w_s[r * HP + uk], a0);
      dh_rec += (a0 + a1) + (a2 + a3);
    }
    __syncthreads();
  }
  if (active) {
#pragma unroll
    for (int k = 0; k < HP; ++k)
      if (k < H) atomicAdd(dWhh + (long)j * H + k, dw[k]);
    atomicAdd(dbhh + j, db);
  }
}

}  // namespace

// see include/b200lrcn.h
B2_API int b2_gru_seq_fwd(const float* G, const float* Whh, const float* bhh, float* out, long out_ld, float* saved,
                          int B, int T, int H, int reverse, void* stream) {
  B2_ARG_CHECK(G && Whh && bhh && out && B > 0 && T > 0 && H > 0, "b2_gru_seq_fwd: null pointer or empty");
  B2_ARG_CHECK(H <= 64, "b2_gru_seq_fwd: hidden size %d > 64 is not supported yet", H);
  const int grid = b2_ceil_div(B, NB);
  cudaStream_t st = (cudaStream_t)stream;
  if (H <= 32) {
    const size_t smem = (size_t)(NB * 32 + 2 * NB * 3 * H) * sizeof(float);
    gru_fwd_kernel<32><<<grid, 128, smem, st>>>(G, Whh, bhh, out, out_ld, saved, B, T, H, reverse);
  } else {
    const size_t smem = (size_t)(NB * 64 + 2 * NB * 3 * H) * sizeof(float);
    gru_fwd_kernel<64><<<grid, 256, smem, st>>>(G, Whh, bhh, out, out_ld, saved, B, T, H, reverse);
  }
  B2_LAUNCH_CHECK("gru_fwd_kernel");
  return 0;
}

// dWhh [3H,H] and dbhh [3H] are ACCUMULATED into (caller zeroes them); dG [B,T,3H] (row stride dG_ld) is overwritten.
B2_API int b2_gru_seq_bwd(const float* dout, long dout_ld, const float* out, long out_ld, const float* saved,
                          const float* Whh, float* dG, long dG_ld, float* dWhh, float* dbhh, int B, int T, int H,
                          int reverse, void* stream) {
  B2_ARG_CHECK(dout && out && saved && Whh && dG && dWhh && dbhh && B > 0 && T > 0 && H > 0,
               "b2_gru_seq_bwd: null pointer or empty");
  B2_ARG_CHECK(H <= 64, "b2_gru_seq_bwd: hidden size %d > 64 is not supported yet", H);
  const int grid = b2_ceil_div(B, NB);
  cudaStream_t st = (cudaStream_t)stream;
  if (H <= 32) {
    const size_t smem = (size_t)(3 * 32 * 32 + NB * 3 * 32 + NB * 32) * sizeof(float);
    gru_bwd_kernel<32><<<grid, 128, smem, st>>>(dout, dout_ld, out, out_ld, saved, Whh, dG, dG_ld, dWhh,
                                                                     dbhh, B, T, H, reverse);
  } else {
    static B2PerDeviceOnce attr;
    if (attr.needed()) {
      B2_CUDA_CHECK(cudaFuncSetAttribute(gru_bwd_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
      attr.mark();
    }
    const size_t smem = (size_t)(3 * 64 * 64 + NB * 3 * 64 + NB * 64) * sizeof(float);
    gru_bwd_kernel<64><<<grid, 256, smem, st>>>(dout, dout_ld, out, out_ld, saved, Whh, dG, dG_ld,
                                                                       dWhh, dbhh, B, T, H, reverse);
  }
  B2_LAUNCH_CHECK("gru_bwd_kernel");
  return 0;
}
