// K1 -- fused frame ingest: uint8 HWC decode buffer -> (gather by frame index) -> cv2-exact
// fixed-point bilinear resize -> optional BGR->RGB swap -> /divisor -> CHW fp32 or bf16.
//
// Replaces, per frame, cv2.resize + cv2.cvtColor + `/255.0` + astype(float32) + permute
// (reference: medsos_lrcn/src/loader_data.py:162-163,182,201,112; crime path lrcn/lrcn.py:136-142).
// The resize is OpenCV's INTER_LINEAR 8-bit algorithm restated in integer arithmetic (11-bit
// coefficients, two-pass rounding) so the uint8 result is bit-identical to cv2; the exact 2x2
// decimation case is OpenCV's INTER_AREA shortcut.  HBM-bound: every source byte is touched
// once (through L1/L2 for the 2x2 taps) and every output element written once, coalesced along x.
#include "common.cuh"

namespace {

struct Tap {
  int i0, i1;
  int w0, w1;
};

// OpenCV resize.cpp coefficient for destination index d (see oracle/lrcn_oracle.py::_linear_coeffs)
__device__ __forceinline__ Tap make_tap(int d, int src, double scale, bool vertical) {
  float f = (float)(((double)d + 0.5) * scale - 0.5);
  int s = (int)floorf(f);
  f -= (float)s;
  Tap t;
  if (vertical) {
    t.i0 = min(max(s, 0), src - 1);
    t.i1 = min(max(s + 1, 0), src - 1);
  } else {
    if (s < 0) {
      f = 0.f;
      s = 0;
    }
    if (s >= src - 1) {
      f = 0.f;
      s = src - 1;
    }
    t.i0 = s;
    t.i1 = min(s + 1, src - 1);
  }
  t.w1 = __float2int_rn(f * 2048.f);
  t.w0 = __float2int_rn((1.f - f) * 2048.f);
  return t;
}

template <typename OutT>
__device__ __forceinline__ OutT to_out(float v);
template <>
__device__ __forceinline__ float to_out<float>(float v) { return v; }
template <>
__device__ __forceinline__ bf16 to_out<bf16>(float v) { return __float2bfloat16_rn(v); }

// mode 0: identity, 1: 2x2 area, 2: general bilinear
template <typename OutT>
__global__ void __launch_bounds__(256)
ingest_kernel(const uint8_t* __restrict__ src, long src_frame_stride, int src_h, int src_w,
              const int* __restrict__ frame_index, int n_out, OutT* __restrict__ dst, int dst_h, int dst_w,
              int swap_rb, float divisor, int mode, double scale_x, double scale_y) {
  const long total = (long)n_out * dst_h * dst_w;
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int x = (int)(idx % dst_w);
    const int y = (int)((idx / dst_w) % dst_h);
    const int f = (int)(idx / ((long)dst_w * dst_h));
    const int sf = frame_index ? frame_index[f] : f;
    int px[3] = {0, 0, 0};
    if (sf >= 0) {
      const uint8_t* fr = src + (long)sf * src_frame_stride;
      if (mode == 0) {
        const uint8_t* p = fr + ((long)y * src_w + x) * 3;
        px[0] = p[0];
        px[1] = p[1];
        px[2] = p[2];
      } else if (mode == 1) {
        const uint8_t* p0 = fr + ((long)(2 * y) * src_w + 2 * x) * 3;
        const uint8_t* p1 = p0 + (long)src_w * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) px[c] = (p0[c] + p0[3 + c] + p1[c] + p1[3 + c] + 2) >> 2;
      } else {
        const Tap tx = make_tap(x, src_w, scale_x, false);
        const Tap ty = make_tap(y, src_h, scale_y, true);
        const uint8_t* r0 = fr + (long)ty.i0 * src_w * 3;
        const uint8_t* r1 = fr + (long)ty.i1 * src_w * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const int h0 = r0[tx.i0 * 3 + c] * tx.w0 + r0[tx.i1 * 3 + c] * tx.w1;
          const int h1 = r1[tx.i0 * 3 + c] * tx.w0 + r1[tx.i1 * 3 + c] * tx.w1;
          int v = (((ty.w0 * (h0 >> 4)) >> 16) + ((ty.w1 * (h1 >> 4)) >> 16) + 2) >> 2;
          px[c] = min(max(v, 0), 255);
        }
      }
    }
    const long plane = (long)dst_h * dst_w;
    OutT* o = dst + (long)f * 3 * plane + (long)y * dst_w + x;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int sc = swap_rb ? 2 - c : c;
      const float v = (divisor == 1.0f) ? (float)px[sc] : __fdiv_rn((float)px[sc], divisor);
      o[c * plane] = to_out<OutT>(v);
    }
  }
}

}  // namespace

B2_API int b2_ingest_u8(const void* src, int n_src_frames, int src_h, int src_w, long src_frame_stride,
                        const int* frame_index, int n_out, void* dst, int dst_h, int dst_w, int out_bf16,
                        int swap_rb, float divisor, void* stream) {
  B2_ARG_CHECK(src && dst, "b2_ingest_u8: null pointer");
  B2_ARG_CHECK(n_src_frames > 0 && src_h > 0 && src_w > 0 && dst_h > 0 && dst_w > 0 && n_out >= 0,
               "b2_ingest_u8: bad shape");
  B2_ARG_CHECK(src_frame_stride >= (long)src_h * src_w * 3, "b2_ingest_u8: frame stride smaller than a frame");
  B2_ARG_CHECK(divisor > 0.f, "b2_ingest_u8: divisor must be positive");
  if (n_out == 0) return 0;
  int mode = 2;
  if (src_h == dst_h && src_w == dst_w) mode = 0;
  else if (src_h == 2 * dst_h && src_w == 2 * dst_w) mode = 1;
  const double sx = (double)src_w / dst_w, sy = (double)src_h / dst_h;
  const long total = (long)n_out * dst_h * dst_w;
  const int grid = (int)((total + 255) / 256 < (long)b2_num_sms() * 16 ? (total + 255) / 256 : (long)b2_num_sms() * 16);
  cudaStream_t st = (cudaStream_t)stream;
  if (out_bf16)
    ingest_kernel<bf16><<<grid, 256, 0, st>>>((const uint8_t*)src, src_frame_stride, src_h, src_w, frame_index, n_out,
                                              (bf16*)dst, dst_h, dst_w, swap_rb, divisor, mode, sx, sy);
  else
    ingest_kernel<float><<<grid, 256, 0, st>>>((const uint8_t*)src, src_frame_stride, src_h, src_w, frame_index,
                                               n_out, (float*)dst, dst_h, dst_w, swap_rb, divisor, mode, sx, sy);
  B2_LAUNCH_CHECK("ingest_kernel");
  return 0;
}
