/* b200lrcn.h -- C ABI of libb200lrcn.so: hand-written sm_100a kernels for the LRCN hot path.
 *
 * The reference (AhmadRifqi86/video-classif) is pure Python/PyTorch and has no FFI of its own;
 * its seam for this path is `nn.Module.forward(clips[B,T,C,H,W]) -> logits` plus autograd
 * (medsos_lrcn/src/models.py:188-234, lrcn/ucf50-lrcn.py:304-336, lrcn/lrcn.py:285-305,
 * lrcn/rgb_lrcn.py:247-263, notebook LRCN nb:174-193).  Every torch call site on that path
 *  maps to one entry point below; the Python host side
 * (video-classif_b200/*.py) binds them with ctypes and wraps them in torch.autograd.Functions.
 *
 * Conventions
 *   - plain pointers are DEVICE pointers unless stated otherwise; no torch types, no ownership
 *     transfer, no allocation inside the library; `stream` is a cudaStream_t passed as void*.
 *   - return value: 0 ok, < 0 argument / environment error, > 0 a cudaError_t.  The message of
 *     the last failure on the calling thread is b2_last_error().
 *   - there is NO CPU or other-GPU fallback: b2_device_check() fails unless the device is sm_100.
 *   - "ACCUMULATED" outputs are added into and must be zeroed by the caller.
 *   - bf16 = 16-bit bfloat, "NHWC" tensors are [N][H][W][C] contiguous, "NCHW" [N][C][H][W].
 */
#ifndef B200LRCN_H_
#define B200LRCN_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- library status ------------------------------------------------------------------ */
int b2_abi_version(void);
const char* b2_last_error(void);
long b2_launch_count(void);
int b2_add_launch_count(long n);   /* kernels replayed from a captured CUDA graph (host layer bookkeeping) */      /* kernels launched by this library since load (all streams) */
int b2_device_check(void);

/* ---- K1 frame ingest -------------------------------------------------------------------
 * Replaces cv2.resize(INTER_LINEAR) + cv2.cvtColor(BGR2RGB) + `/255.0` + float32 + permute
 * (medsos_lrcn/src/loader_data.py:162-163,182,201,112; crime path lrcn/lrcn.py:136-142) and the
 * frame gather of uniform_sampling/duplicate_frames (loader_data.py:35-51).
 * src: uint8 [n_src_frames][src_h][src_w][3] (frame stride in bytes); frame_index: n_out source
 * frame numbers (device int32, NULL = identity, -1 = all-zero frame, lrcn/lrcn.py:155);
 * dst: [n_out][3][dst_h][dst_w] fp32 or bf16.  The uint8 resize result is bit-identical to cv2. */
int b2_ingest_u8(const void* src, int n_src_frames, int src_h, int src_w, long src_frame_stride,
                 const int* frame_index, int n_out, void* dst, int dst_h, int dst_w, int out_bf16,
                 int swap_rb, float divisor, void* stream);

/* ---- K2 tensor-core GEMM / implicit-GEMM convolution (tcgen05 + TMEM + TMA) -------------
 * b2_gemm_bf16_tn: D[M,N] = A[M,K] B[N,K]^T (+bias[N]) (ReLU); A,B bf16 row-major (lda/ldb in
 * elements, multiples of 8), D bf16 or fp32 (ldd elements).  Replaces nn.Linear / the hoisted
 * nn.LSTM input GEMM / 1x1 convolutions (models.py:200-202,213,221-226).
 * col_sum/col_sumsq (fp32 [N], ACCUMULATED, may be NULL): per-column sum and sum of squares of
 * the stored output -- the batch statistics of the BatchNorm that follows a convolution.
 * b2_conv2d_nhwc_bf16: y[N,P,Q,Cout] = conv(x[N,H,W,C], w[Cout][R][S][C]); C % 64 == 0; zero
 * padding `pad`, stride `stride`; replaces torchvision ResNet Conv2d layers (models.py:192). */
int b2_gemm_bf16_tn(const void* A, long lda, const void* B, long ldb, void* D, long ldd, int M, int N,
                    int K, const float* bias, const float* bias2, int out_bf16, int relu, float* col_sum,
                    float* col_sumsq, void* stream);
int b2_conv2d_nhwc_bf16(const void* x, int N, int H, int W, int C, const void* w, int Cout, int R, int S,
                        int stride, int pad, void* y, const float* bias, int out_bf16, int relu,
                        float* col_sum, float* col_sumsq, void* stream);

/* Convolution with the train-mode BatchNorms of a ResNet block folded in (torchvision Bottleneck /
 * BasicBlock under models.py:192, train_eval.py:12), all optional (NULL = off):
 *   a_scale/a_shift [C]   : the INPUT is a raw conv output; relu?(x*scale+shift) is applied to the A
 *                           operand tile in shared memory (zero padding stays zero) -- the normalised
 *                           tensor is never materialised
 *   col_sum/col_sumsq [Cout] (ACCUMULATED): statistics of the raw output; y = NULL makes this a
 *                           statistics-only pass (nothing is written)
 *   fin_*                 : the last CTA converts (col_sum, col_sumsq) into fin_scale/fin_shift [Cout]
 *                           (biased variance, eps) and updates the running stats (momentum, unbiased
 *                           variance); *fin_counter must be 0 at launch
 *   o_scale/o_shift [Cout]: BatchNorm of the OUTPUT applied in the epilogue on the fp32 accumulators
 *   res [N,P,Q,Cout] bf16 : shortcut added before the final ReLU; r_scale/r_shift = its own BatchNorm
 *                           (down-sample branch stored raw) or NULL (identity shortcut)
 * Bottleneck: conv1 (stats+fin) -> conv2 (a=bn1, stats+fin) -> conv3 statistics pass (a=bn2, y=NULL,
 * stats+fin) -> conv3 output pass (a=bn2, o=bn3, res, relu). */
int b2_conv2d_bn_nhwc_bf16(const void* x, int N, int H, int W, int C, const void* w, int Cout, int R, int S,
                           int stride, int pad, void* y, const float* a_scale, const float* a_shift, int a_relu,
                           const float* o_scale, const float* o_shift, const void* res, const float* r_scale,
                           const float* r_shift, int relu, float* col_sum, float* col_sumsq,
                           const float* fin_gamma, const float* fin_beta, float* fin_running_mean,
                           float* fin_running_var, float* fin_scale, float* fin_shift, unsigned int* fin_counter,
                           float eps, float momentum, void* stream);
/* scale/shift of one BatchNorm2d from batch statistics (train: also the running-stat update) or from the
 * running statistics (train = 0); b2_scale_shift_apply_nhwc: y = act(x*scale+shift [+ res | + res*rscale+rshift]). */
int b2_bn_finalize_nhwc(const float* sum, const float* sumsq, const float* gamma, const float* beta,
                        float* running_mean, float* running_var, long count, float eps, float momentum, int train,
                        float* scale, float* shift, int C, void* stream);
int b2_scale_shift_apply_nhwc(const void* x, void* y, long rows, int C, const float* scale, const float* shift,
                              const void* res, const float* rscale, const float* rshift, int relu, void* stream);

/* 3x3 / stride 1 / pad 1 convolution of a narrow stage (C = Cout = 64: Bottleneck.conv2 of layer1, BasicBlock convs
 * of layer1) as a halo-tile kernel: one TMA brings the (TH+2) x (W+2) input halo of TH output rows into shared
 * memory and the nine filter taps are nine shifted descriptors over that tile (no 9x re-fetch through L2), the 72 KB
 * of weights stay resident; a_scale/a_shift (optional) = BatchNorm(+ReLU) of the INPUT applied to the halo once per
 * tile; col_sum/col_sumsq, fin_* as b2_conv2d_bn_nhwc_bf16.  b2_conv3x3_halo_supported() tells whether a shape is
 * covered (otherwise use b2_conv2d_bn_nhwc_bf16). */
int b2_conv3x3_halo_supported(int N, int H, int W, int C, int Cout);
int b2_conv3x3_halo_bn_nhwc_bf16(const void* x, int N, int H, int W, int C, const void* w, int Cout, void* y,
                                 const float* a_scale, const float* a_shift, int a_relu, float* col_sum, float* col_sumsq,
                                 const float* fin_gamma, const float* fin_beta, float* fin_running_mean,
                                 float* fin_running_var, float* fin_scale, float* fin_shift, unsigned int* fin_counter,
                                 float eps, float momentum, void* stream);

/* BatchNorm statistics + finalisation of a 1x1 convolution y = relu?(x*a_scale+a_shift) W^T WITHOUT computing y
 * (torchvision Bottleneck.bn3 in train mode, models.py:192): per-channel sum_m y and sum_m y^2 follow from the
 * K-vector s = sum_m t[m,:] and the K x K Gram matrix G = sum_m t^T t of the transformed input (tensor-core
 * MN-major MMA over one streaming pass of x), evaluated per output channel in fp64.  x [M, C] bf16 (C = 64, 128 or 256),
 * w [Cout, C] bf16; workspace: b2_gram_workspace_floats(C) floats (scratch, zeroed inside); col_sum/col_sumsq [Cout] optional outputs
 * (OVERWRITTEN); fin_scale/fin_shift [Cout] = gamma*rstd, beta - mean*gamma*rstd; running stats updated when given. */
long b2_gram_workspace_floats(int C);
int b2_conv1x1_gram_bnstats_bf16(const void* x, long M, int C, const void* w, int Cout, const float* a_scale,
                                 const float* a_shift, int a_relu, float* workspace, float* col_sum, float* col_sumsq,
                                 const float* fin_gamma, const float* fin_beta, float* fin_running_mean,
                                 float* fin_running_var, float* fin_scale, float* fin_shift, float eps, float momentum,
                                 void* stream);

/* ---- backbone glue (NHWC bf16) ----------------------------------------------------------
 * b2_stem_im2col: NCHW fp32/bf16 frames -> [N*P*Q][Kp] bf16 patches of the 7x7/2 pad-3 stem conv,
 * column k = (c*7+r)*8+s (filter rows padded to 8 taps), Kp = 168.
 * b2_bn_apply_nhwc: y = act( BN(x) [+ res | + BN2(res)] ), BatchNorm2d semantics of torch
 * (train: batch statistics from sum/sumsq over `count` elements, running stats updated with
 * momentum and unbiased variance; eval: running stats) -- models.py:192 under train_eval.py:12.
 * res_mode 0 none, 1 identity shortcut (already activated), 2 down-sample branch (raw + own BN).
 * b2_bn_relu_maxpool_nhwc: stem BN + ReLU + MaxPool2d(3,2,1).  b2_avgpool_nhwc: AdaptiveAvgPool2d(1). */
int b2_stem_im2col(const void* x, int in_bf16, void* A, int N, int H, int W, int Kp, void* stream);
/* Direct stem convolution (Conv2d 3->64, 7x7, stride 2, pad 3; torchvision ResNet.conv1 under
 * models.py:192) without a patch matrix: b2_stem_pack repacks NCHW fp32/bf16 frames into zero-padded
 * bf16 units of 2 pixels x 4 channels, even/odd rows in separate planes (b2_stem_packed_bytes bytes);
 * b2_stem_conv_bf16 runs the convolution as a Toeplitz-descriptor tcgen05 GEMM over that stream and
 * writes raw NHWC bf16 [N][P][Q][64] + the per-channel sum / sum of squares (ACCUMULATED, may be NULL).
 * wk: weights as bf16 [28][64][8] with k = r*32 + s*4 + c (s = 7 and c = 3 slots zero). */
long b2_stem_packed_bytes(int N, int H, int W);
int b2_stem_pack(const void* x, int in_bf16, void* xp, int N, int H, int W, void* stream);
int b2_stem_conv_bf16(const void* xp, const void* wk, void* y, int N, int H, int W, float* col_sum,
                      float* col_sumsq, void* stream);
int b2_bn_apply_nhwc(const void* x, void* y, long rows, int C, const float* sum, const float* sumsq,
                     const float* gamma, const float* beta, float* running_mean, float* running_var,
                     int res_mode, const void* res, const float* rsum, const float* rsumsq,
                     const float* rgamma, const float* rbeta, float* rrunning_mean, float* rrunning_var,
                     long count, float eps, float momentum, int train, int relu, void* stream);
int b2_bn_relu_maxpool_nhwc(const void* x, void* y, int N, int H, int W, int C, const float* sum,
                            const float* sumsq, const float* gamma, const float* beta, float* running_mean,
                            float* running_var, float eps, float momentum, int train, void* stream);
int b2_avgpool_nhwc(const void* x, float* out_f32, void* out_bf16, int N, int HW, int C, void* stream);

/* ---- trainable tail ---------------------------------------------------------------------
 * b2_act_ln_fwd/bwd: out = LayerNorm(act(pre)) with act = exact-erf GELU (apply_gelu=1) or
 * identity -- `self.bnK(F.gelu(self.adaptK(x)))` models.py:200-202, `self.bn0(rnn_out)` :222.
 * b2_sgemm: C = alpha op(A) op(B) + beta C, fp32 row-major (fp32 parity path, tiny-N layers,
 * weight/input gradients).  b2_colsum_f32: bias gradients.  casts feed the bf16 GEMM. */
int b2_act_ln_fwd(const float* pre, const float* gamma, const float* beta, float* out_f32, void* out_bf16,
                  float* mean, float* rstd, long M, int N, float eps, int apply_gelu, void* stream);
int b2_act_ln_bwd(const float* dout, const float* pre, const float* gamma, const float* mean,
                  const float* rstd, float* dpre, float* dgamma /*ACCUMULATED*/, float* dbeta /*ACCUMULATED*/,
                  long M, int N, int apply_gelu, void* stream);
int b2_sgemm(int trans_a, int trans_b, int M, int N, int K, float alpha, const float* A, long lda,
             const float* B, long ldb, float beta, float* C, long ldc, const float* bias, const float* bias2,
             void* stream);
int b2_colsum_f32(const float* X, long ld, long M, int N, float* out, int accumulate, void* stream);
int b2_cast_f32_bf16(const float* src, void* dst, long n, void* stream);
int b2_transpose_cast_f32_bf16(const float* src, long ld_src, void* dst, long ld_dst, long R, long C,
                               void* stream);
/* nn.Dropout(p) in train mode (nb:163, models.py:153,182): y = x*keep/(1-p) with a counter-hash
 * mask; calling it again with the same seed on dy replays the mask for the backward pass. */
int b2_dropout_f32(const float* x, float* y, long n, float p, unsigned long long seed, const unsigned long long* seed_offset,
                   void* stream);   /* seed_offset (device, may be NULL) is added to seed inside the kernel: a launch captured in a
                                       CUDA graph draws a new mask per replay when the caller bumps that counter */
/* stand-alone activations (models_bidir.py:119-155 Adapt 's'/'g'/'r', F.silu head): kind 0 relu, 1 gelu (erf), 2 silu; x = the pre-activation */
int b2_act_fwd_f32(const float* x, float* y, long n, int kind, void* stream);
int b2_act_bwd_f32(const float* dy, const float* x, float* dx, long n, int kind, void* stream);
int b2_transpose_bf16(const void* src, long ld_src, void* dst, long ld_dst, long R, long C, void* stream);   /* dst[c][r] = src[r][c], bf16 */

/* ---- K3/K4 persistent LSTM --------------------------------------------------------------
 * torch.nn.LSTM(batch_first=True) semantics (nb:169, models.py:156-158, lrcn.py:236): gate order
 * i,f,g,o, zero initial state.  G[B][T][4H] = x W_ih^T + b_ih + b_hh is computed beforehand by the
 * GEMM; out points at column dir*H of the [B][T][dirs*H] layer output (out_ld = dirs*H).
 * gates/cstate ([B][T][4H] / [B][T][H], post-activation) are saved for BPTT (NULL at inference).
 * b2_lstm_seq_bwd writes dG[B*T][4H] (row stride dG_ld) and ACCUMULATES dWhh[4H][H].
 * bias/bias2 of b2_gemm_bf16_tn / b2_sgemm are added per output column (b_ih + b_hh). */
int b2_lstm_seq_fwd(const float* G, const float* Whh, float* out, long out_ld, float* gates, float* cstate,
                    int B, int T, int H, int reverse, void* stream);
int b2_lstm_seq_bwd(const float* dout, long dout_ld, const float* out, long out_ld, const float* gates,
                    const float* cstate, const float* Whh, float* dG, long dG_ld, float* dWhh, int B, int T,
                    int H, int reverse, void* stream);

/* ---- small TimeDistributed CNN, fp32 NCHW (nb:156-163,181-183; backup_ucf50.py:113-140) --
 * b2_conv3x3_f32: Conv2d(k=3,padding=1) forward (transposed=0, w [Cout][Cin][3][3]) or data
 * gradient (transposed=1: pass dy as x, Cin=Cout_fwd, Cout=Cin_fwd, same w).
 * b2_conv3x3_wgrad_f32: ACCUMULATES dw.  b2_bn2d_*: train-mode BatchNorm2d statistics (double
 * accumulators, ACCUMULATED), finalisation (scale/shift/mean/rstd + running-stat update), fused
 * BN+ReLU(+MaxPool2d(2,2)) forward and its two-pass backward (s1 = dbeta, s2 = dgamma). */
int b2_conv3x3_f32(const float* x, const float* w, const float* bias, float* y, int N, int Cin, int Cout,
                   int H, int W, int transposed, void* stream);
int b2_conv3x3_wgrad_f32(const float* x, const float* dy, float* dw, int N, int Cin, int Cout, int H, int W,
                         void* stream);
int b2_bn2d_stats_f32(const float* x, int N, int C, int HW, double* sum, double* sumsq, void* stream);
int b2_bn2d_finalize(const double* sum, const double* sumsq, long count, const float* gamma,
                     const float* beta, float* running_mean, float* running_var, float momentum, float eps,
                     int train, float* scale, float* shift, float* mean, float* rstd, int C, void* stream);
int b2_bn2d_act_pool_fwd_f32(const float* x, const float* scale, const float* shift, float* y, void* y_bf16,
                             int N, int C, int H, int W, int pool, void* stream);
int b2_bn2d_act_pool_bwd_reduce_f32(const float* x, const float* dy, const float* scale, const float* shift,
                                    const float* mean, const float* rstd, int N, int C, int H, int W,
                                    int pool, double* s1, double* s2, void* stream);
int b2_bn2d_act_pool_bwd_apply_f32(const float* x, const float* dy, const float* scale, const float* shift,
                                   const float* mean, const float* rstd, const float* gamma, const double* s1,
                                   const double* s2, long count, int train, float* dx, int N, int C, int H,
                                   int W, int pool, void* stream);
int b2_f64_to_f32(const double* src, float* dst, int n, int accumulate, void* stream);

/* ---- small-CNN tensor-core path (csrc/smallcnn_tc.cu): the notebook CNN (nb:156-163,181-183; backup_ucf50.py:113-115,138-140)
 * in bf16 NHWC on tcgen05.  One pixel = one swizzle row of 2*C bytes (C = 16 / 32 / 64 -> SWIZZLE_32B / 64B / 128B); a tile is
 * TH whole output rows, its halo comes in by one TMA and every filter tap reads it through a shifted descriptor.
 * b2_sc_conv3x3_bf16: y [N,H,W,Cout] = conv3x3(x [N,H,W,Cin], w [Cout][3][3][Cin]) (+ bias), stride 1, pad 1, + optional
 *   per-channel sum / sum of squares of the stored values (ACCUMULATED).  (Cin, Cout) in {(16,16), (16,32), (32,64)} (forward) and
 *   {(64,32), (32,16)} (their data gradients: w = the flipped, transposed filter).
 * b2_sc_conv3x3_wgrad_bf16: dw [Cout][3][3][Cin] fp32 (ACCUMULATED) = sum over pixels of dz [N,H,W,Cout] x shifted x [N,H,W,Cin].
 * b2_sc_conv1_fwd / _wgrad: the 3-channel first layer on CUDA cores (x fp32 NCHW [N,3,H,W], w / dw fp32 [16][3][3][3] in torch
 *   layout, y / dz bf16 NHWC [N,H,W,16]); dw ACCUMULATED.
 * b2_sc_pack_input: x fp32 NCHW [N,3,H,W] -> bf16 NHWC [N,H,W,16] (channels 3..15 zero) so that the first layer runs on the same
 *   tensor-core kernels ((Cin, Cout) = (16,16), filter zero-padded to 16 input channels).
 * b2_sc_bn_finalize: (sum, sumsq, count) of the BIAS-FREE raw conv output + the conv bias -> scale, shift, mean, rstd for that
 *   tensor (+ running statistics of raw + bias, momentum; train = 0: from the running statistics).
 * b2_sc_act_pool_fwd: y = maxpool_pool(relu(raw * scale + shift)), pool in {1, 2}; _bwd_reduce: s1 = sum dpre (= dbeta),
 *   s2 = sum dpre * xhat (= dgamma), ACCUMULATED, dpre = dy routed to the first maximum of its window where positive;
 *   _bwd_apply: dz [N,H,W,C] = scale * (dpre - s1/M - xhat s2/M) (train) or scale * dpre (eval).
 * b2_sc_nhwc_to_chw: feat [N][C*HW] bf16 (the channel-major flatten of nb:186) = dropout_p(act [N][HW][C]); b2_sc_chw_to_nhwc:
 *   the inverse on the fp32 / bf16 gradient with the same mask (same seed). */
int b2_sc_conv3x3_bf16(const void* x, int N, int H, int W, int Cin, const void* w, int Cout, void* y, const float* bias,
                       float* col_sum, float* col_sumsq, void* stream);
int b2_sc_conv3x3_wgrad_bf16(const void* x, const void* dz, int N, int H, int W, int Cin, int Cout, float* dw, void* stream);
int b2_sc_conv1_fwd(const float* x, const float* w, const float* bias, void* y, int N, int H, int W, float* col_sum,
                    float* col_sumsq, void* stream);
int b2_sc_conv1_wgrad(const float* x, const void* dz, float* dw, int N, int H, int W, void* stream);
int b2_sc_pack_input(const float* x, void* y, int N, int H, int W, void* stream);
int b2_sc_bn_finalize(const float* sum, const float* sumsq, const float* conv_bias, const float* gamma, const float* beta,
                      float* running_mean, float* running_var, long count, float eps, float momentum, int train, float* scale,
                      float* shift, float* mean, float* rstd, int C, void* stream);
int b2_sc_act_pool_fwd(const void* raw, const float* scale, const float* shift, void* y, int N, int H, int W, int C, int pool,
                       void* stream);
int b2_sc_act_pool_bwd_reduce(const void* raw, const void* dy, const float* scale, const float* shift, const float* mean,
                              const float* rstd, float* s1, float* s2, int N, int H, int W, int C, int pool, void* stream);
int b2_sc_act_pool_bwd_apply(const void* raw, const void* dy, const float* scale, const float* shift, const float* mean,
                             const float* rstd, float* s1, float* s2, int train, void* dz, int N, int H, int W, int C, int pool,
                             void* stream);
int b2_sc_nhwc_to_chw(const void* act, void* feat, int N, int HW, int C, float p_drop, unsigned long long seed,
                      const unsigned long long* seed_offset, void* stream);
int b2_sc_chw_to_nhwc(const void* dfeat, int in_bf16, void* dact, int N, int HW, int C, float p_drop, unsigned long long seed,
                      const unsigned long long* seed_offset, void* stream);

/* ---- persistent GRU layer (torch.nn.GRU semantics, gate order r,z,n; lrcn/backup_ucf50.py:126, medsos models.py:160-170)
 * G [B,T,3H] = x W_ih^T + b_ih (hoisted gate GEMM); Whh [3H,H]; bhh [3H] (b_hn stays inside the r product);
 * out [B,T,*] with row stride out_ld (a direction's slice of the [B,T,dirs*H] output); saved [B,T,4H] = r,z,n,hh_n for
 * BPTT (NULL for inference); reverse = 1 runs t = T-1..0 (the `_reverse` parameters).
 * Backward: dG [B,T,3H] (row stride dG_ld) = gradient of G; dWhh / dbhh ACCUMULATED into (caller zeroes). */
int b2_gru_seq_fwd(const float* G, const float* Whh, const float* bhh, float* out, long out_ld, float* saved, int B, int T,
                   int H, int reverse, void* stream);
int b2_gru_seq_bwd(const float* dout, long dout_ld, const float* out, long out_ld, const float* saved, const float* Whh,
                   float* dG, long dG_ld, float* dWhh, float* dbhh, int B, int T, int H, int reverse, void* stream);

/* ---- whole-stack persistent LSTM (unidirectional nn.LSTM, H <= 64, input width <= 64, T <= 64, <= 8 layers):
 * every layer and timestep in one launch; medsos_lrcn/src/models.py:156-158,205 as tuned by all_config.py:14-17.
 * w_ih / w_hh / b_ih / b_hh (and dw_ih / dw_hh / db) are HOST arrays of `layers` device pointers in nn.LSTM layout
 * (weight_ih_l{k} [4H, in_k], weight_hh_l{k} [4H, H], gate order i,f,g,o).  out [layers, B, T, H] holds every
 * layer's output sequence (the module output is out[layers-1]); gates [layers, B, T, 4H] / cstate [layers, B, T, H]
 * are saved for BPTT (NULL for inference).  Backward: dout [B, T, H] = gradient of the top layer's output sequence;
 * dw_ih / dw_hh / db are ACCUMULATED into (caller zeroes); db[k] is the gradient of b_ih_l{k} and of b_hh_l{k};
 * dx [B, T, In0] may be NULL. */
int b2_lstm_stack_fwd(const float* x, int In0, const void* const* w_ih, const void* const* w_hh, const void* const* b_ih,
                      const void* const* b_hh, int layers, float* out, float* gates, float* cstate, int B, int T, int H,
                      void* stream);
int b2_lstm_stack_bwd(const float* dout, const float* x, int In0, const void* const* w_ih, const void* const* w_hh,
                      int layers, const float* out, const float* gates, const float* cstate, float* dx,
                      void* const* dw_ih, void* const* dw_hh, void* const* db, int B, int T, int H, void* stream);

/* ---- selective scan forward (VideoMamba temporal mixer; lrcn/videomamba.py:242-284 parallel_scan,
 * medsos_lrcn/src/models.py:47-71) -------------------------------------------------------------------
 * x_t = exp(delta_t A) x_{t-1} + delta_t B_t u_t ; y_t = <x_t, C_t>.  u, delta, y [batch, L, D] fp32;
 * A [D, N]; B, C [batch, L, N]; any N in 1..64 (the medsos search grid gives n_state = hidden in {12,...,64}: the kernels
 * run on the next compiled width b2_scan_padded_states(N) in {4, 8, 16, 32, 64} with the padding states held at zero).
 * chunk_reset > 0: the state restarts from zero every chunk_reset steps (videomamba.py resets per 256-step
 * chunk); <= 0: one scan over L.  reverse = 1: u and delta are read time-reversed, B and C are not, y is written
 * time-reversed (the medsos "backward" direction). */
int b2_selective_scan_fwd(const float* u, const float* delta, const float* A, const float* B, const float* C, float* y,
                          int batch, int L, int D, int N, int chunk_reset, int reverse, void* stream);
int b2_scan_padded_states(int N);

/* Elementwise pieces of the Mamba ResidualBlock around the scan (medsos_lrcn/src/models.py:9-117), forward only, fp32:
 * RMSNorm over the last dim; causal depthwise Conv1d over time (k taps, padding k-1, output trimmed to L) + SiLU on
 * [B, L, D] tensors (x row stride x_ld >= D: a column slice of in_proj's output); softplus (threshold 20);
 * y = a * silu(res[:, c % res_cols]) (res row stride res_ld). */
int b2_rmsnorm_f32(const float* x, const float* w, float* y, long rows, int D, float eps, void* stream);
int b2_dwconv1d_silu_f32(const float* x, long x_ld, const float* w, const float* b, float* y, int B, int L, int D, int K,
                         void* stream);
int b2_softplus_f32(const float* x, float* y, long n, void* stream);
int b2_mul_silu_f32(const float* a, const float* res, long res_ld, int res_cols, float* y, long rows, int cols, void* stream);

/* Backward of the same pieces (training rnn_type="mamba": small L / D / N).  Buffers marked ACCUMULATED are added into
 * with atomics and must be zeroed by the caller.  b2_selective_scan_bwd: scans / chunks of up to 512 steps keep the recomputed forward
 * states on chip (checkpoints in shared memory, segments in registers) and take workspace = NULL; longer ones need
 * workspace[b2_scan_bwd_workspace_floats(...)] (= batch*D*L*b2_scan_padded_states(N), 0 when none is needed); chunk_reset > 0: state reset every chunk_reset steps (videomamba), chunks in parallel; a_is_log = 1: dA is the
 * gradient of A_log where A = -exp(A_log), 0: of A itself. */
int b2_rmsnorm_bwd_f32(const float* dy, const float* x, const float* w, float* dx, float* dw, long rows, int D, float eps,
                       void* stream);
int b2_dwconv1d_silu_bwd_f32(const float* dy, const float* x, long x_ld, const float* w, const float* b, float* dx, long dx_ld,
                             float* dw, float* db, int B, int L, int D, int K, void* stream);
int b2_softplus_bwd_f32(const float* dy, const float* x, float* dx, long n, void* stream);
int b2_mul_silu_bwd_f32(const float* dy, const float* a, const float* res, long res_ld, int res_cols, float* da, float* dres,
                        long dres_ld, long rows, int cols, void* stream);
long b2_scan_bwd_workspace_floats(int batch, int L, int D, int N, int chunk_reset);
int b2_selective_scan_bwd(const float* u, const float* delta, const float* A, const float* B, const float* C, const float* dy,
                          float* workspace, float* du, float* ddelta, float* dA, float* dB, float* dC, int batch, int L,
                          int D, int N, int chunk_reset, int reverse, int a_is_log, void* stream);

/* ---- trainable frame encoder: backward kernels (csrc/conv_bwd.cu; fine-tune paths rgb_lrcn.py:208-245, lrcn.py:246-283) ----
 * b2_conv2d_wgrad_nhwc_bf16: dw[Cout,R,S,C] fp32 += sum_m dy[m,co] * x[pix(m)+(r,s),ci] (tcgen05, MN-major operands straight
 *   from the NHWC tensors; ACCUMULATED: the caller zeroes dw).  R = S = 1, stride 1: C % 8 == 0; otherwise C % 64 == 0.
 * b2_bn_bwd_nhwc_bf16: BatchNorm backward over [M,C] bf16: dzm = dz * [z > 0] when z != NULL (else dz is used); s1 (= dbeta), s2
 *   (= dgamma) ACCUMULATED; dy = gamma*invstd*(dz - s1/count - xhat*s2/count) (train) or gamma*invstd*dz (eval).
 * b2_dilate2_nhwc_bf16: zero-dilation of dy for the data gradient of stride-2 convs (z pre-zeroed).
 * b2_avgpool_bwd_nhwc: dz[n,hw,c] = dfeat[n,c] / HW.   b2_maxpool_relu_bwd_nhwc: stem tail backward (dbn fp32 pre-zeroed). */
/* b2_conv_weight_layouts: torch conv weight w [Cout,Cin,R,S] fp32 -> wk [Cout,R,S,Cin] bf16 (forward / weight-gradient layout)
 * and wt [Cin,R,S,Cout] bf16 with flipped taps (data-gradient layout) in one pass; either output may be NULL. */
int b2_conv_weight_layouts(const float* w, void* wk, void* wt, int Cout, int Cin, int R, int S, void* stream);
int b2_conv2d_wgrad_nhwc_bf16(const void* x, int Nimg, int H, int W, int C, const void* dy, int Cout, int R, int S, int stride,
                              int pad, float* dw, void* stream);
int b2_bn_bwd_nhwc_bf16(const void* dz, void* dzm, const void* z, const void* y, void* dy, const float* gamma, const float* sum,
                        const float* sumsq, const float* running_mean, const float* running_var, float* s1, float* s2, long M,
                        int C, long count, float eps, int train, void* stream);
int b2_dilate2_nhwc_bf16(const void* dy, void* z, int N, int P, int Q, int H, int W, int C, void* stream);
int b2_avgpool_bwd_nhwc(const float* dfeat, void* dz, long N, int HW, int C, void* stream);
int b2_maxpool_relu_bwd_nhwc(const void* raw, const float* scale, const float* shift, const void* dpool, float* dbn, int N,
                             int H, int W, int P, int Q, int C, void* stream);

/* ---- DenseNet frame encoder element kernels (csrc/dense_ops.cu; densenet121 of lrcn/lrcn.py:196-209, rgb_lrcn.py:180-193) ----
 * A dense block is ONE channel-concatenated NHWC buffer (row stride = final channel count); these kernels work on
 * row-strided bf16 tensors.  b2_scale_shift_apply_ld_bf16: y[r,:C] = act(x[r,:C]*scale+shift) (scale = NULL: slice copy);
 * b2_colstats_ld_bf16: per-channel sum / sumsq, ACCUMULATED; b2_avgpool2x2_nhwc_bf16: AvgPool2d(2,2) into a strided dst. */
/* D[M,N] bf16 = relu?(A[M,K]*a_scale[k]+a_shift[k]) B[N,K]^T (+ column statistics): the pre-activation BatchNorm + ReLU of a
 * dense layer / transition folded into the 1x1 conv's A-tile transform; A row-strided (lda); a_scale / a_shift readable up
 * to the next multiple of 64 channels (zero padded); N % 32 == 0, N >= 64. */
int b2_gemm_bn_bf16_tn(const void* A, long lda, const void* B, long ldb, void* D, long ldd, int M, int N, int K,
                       const float* a_scale, const float* a_shift, int a_relu, float* col_sum, float* col_sumsq, void* stream);
/* Halo-tile 3x3 conv of a dense layer (128 -> 32 channels, stride 1, pad 1): the input halo is fetched once for all 9 taps,
 * the 32 new channels are written straight into a channel slice of the block buffer (row stride ldy) + their statistics. */
int b2_conv3x3_halo_dense_supported(int N, int H, int W, int C, int Cout);
int b2_conv3x3_halo_dense_bf16(const void* x, int N, int H, int W, int C, const void* w, int Cout, void* y, long ldy,
                               float* col_sum, float* col_sumsq, void* stream);
int b2_scale_shift_apply_ld_bf16(const void* x, long ldx, void* y, long ldy, long rows, int C, const float* scale,
                                 const float* shift, int relu, void* stream);
int b2_colstats_ld_bf16(const void* x, long ld, long rows, int C, float* sum, float* sumsq, void* stream);
int b2_avgpool2x2_nhwc_bf16(const void* x, void* y, long ldy, int N, int H, int W, int C, void* stream);
/* backward twins: BatchNorm (+ReLU mask) backward between row-strided tensors, result written / ACCUMULATED into the
 * block's gradient buffer; AvgPool2d(2,2) backward. */
int b2_bn_bwd_ld_bf16(const void* dz, long lddz, const void* z, long ldz, const void* x, long ldx, void* dx, long lddx,
                      int accumulate, const float* gamma, const float* sum, const float* sumsq, const float* running_mean,
                      const float* running_var, float* s1, float* s2, long M, int C, long count, float eps, int train,
                      void* stream);
int b2_avgpool2x2_bwd_nhwc_bf16(const void* dy, long lddy, void* dx, int N, int H, int W, int C, void* stream);

/* ---- MobileNetV2 frame encoder kernels (csrc/mobilenet_ops.cu; mobilenet_v2 under medsos models.py:133-143, in the search
 * space of medsos_lrcn/src/automation.py:28).  b2_mbv2_stem_conv: Conv2d(3,32,3,s2,p1) NCHW frames -> NHWC bf16 (+ statistics);
 * b2_dwconv3x3_bn_nhwc_bf16: depthwise 3x3 with the previous BatchNorm + activation (act 0 none / 1 ReLU / 2 ReLU6) applied on
 * load, raw output + statistics.  (b2_scale_shift_apply_ld_bf16's `relu` argument takes the same 0 / 1 / 2.) */
int b2_mbv2_stem_conv(const void* x, int in_bf16, const float* w, void* y, float* sum, float* sumsq, int N, int H, int W,
                      void* stream);
int b2_dwconv3x3_bn_nhwc_bf16(const void* x, const float* scale, const float* shift, int act, const float* w, void* y,
                              float* sum, float* sumsq, int N, int H, int W, int C, int stride, void* stream);
/* backward of the trainable MobileNetV2 (lrcn/lrcn.py:196-230,246-283, rgb_lrcn.py:208-227 with CNN_BACKBONE = "mobilenet_v2"):
 * depthwise data / weight gradients (dw fp32 [C,1,3,3] ACCUMULATED), the stem's weight gradient (dw fp32 [32,3,3,3] ACCUMULATED),
 * and BatchNorm backward with the ReLU6 mask 0 < z < 6 (same contract as b2_bn_bwd_nhwc_bf16). */
int b2_dwconv3x3_dgrad_nhwc_bf16(const void* dy, const float* w, void* dx, int N, int H, int W, int C, int stride, void* stream);
int b2_dwconv3x3_wgrad_nhwc_bf16(const void* x, const void* dy, float* dw, int N, int H, int W, int C, int stride, void* stream);
int b2_mbv2_stem_wgrad(const void* x, int in_bf16, const void* dy, float* dw, int N, int H, int W, void* stream);
int b2_bn_bwd_relu6_nhwc_bf16(const void* dz, void* dzm, const void* z, const void* y, void* dy, const float* gamma,
                              const float* sum, const float* sumsq, const float* running_mean, const float* running_var,
                              float* s1, float* s2, long M, int C, long count, float eps, int train, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200LRCN_H_ */
