// K1 -- fused frame ingest: uint8 HWC decode buffer -> (gather by frame index) -> cv2-exact
// fixed-point bilinear resize -> optional BGR->RGB swap -> /divisor -> CHW fp32 or bf16.
//
// Replaces, per frame, cv2.resize + cv2.cvtColor + `/255.0` + astype(float32) + permute
// (reference: medsos_lrcn/src/loader_data.py:162-163,182,201,112; crime path lrcn/lrcn.py:136-142).
// The resize is OpenCV's INTER_LINEAR 8-bit algorithm restated in integer arithmetic (11-bit
// coefficients, two-pass rounding) so the uint8 result is bit-identical to cv2; the exact 2x2
// decimation case is OpenCV's INTER_AREA shortcut.
//   * ingest_identity_vec_kernel  (no resize, W % 16 == 0: the bench's 112x112 clips, cfg 1's 64x64): a thread converts 16 pixels --
//     three 128-bit loads of the interleaved bytes, per channel plane four 128-bit (fp32) or two 128-bit (bf16) stores; byte ->
//     float through a 256-entry table of correctly rounded quotients (the division was the issue-slot hog): 0.30 -> 0.56 of copy peak
//   * ingest_kernel               general resize: one output pixel per thread gathering its 2 x 2 taps through L1 (neighbouring
//     threads share the 32-byte sectors).  It moves (source rows touched + output) at 4.4 TB/s = 0.67 of the copy peak; a
//     row-strip variant that staged the two source rows of every output row in shared memory with 128-bit loads was measured
//     SLOWER (171 vs 135 us at 1024 x 360x640 -> 112x112): the 11-bit fixed-point arithmetic (~150 instructions per pixel)
//     bounds the kernel, not the loads (ncu: 61 % issue-slot utilisation, DRAM 440 MB read in both forms).
#include "common.cuh"

#include <stdlib.h>

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
  __shared__ float lut[256];                     // byte -> float: correctly rounded v / divisor, computed once per CTA
  for (int v = threadIdx.x; v < 256; v += blockDim.x) lut[v] = divisor != 1.0f ? __fdiv_rn((float)v, divisor) : (float)v;
  __syncthreads();
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
      o[c * plane] = to_out<OutT>(lut[px[sc]]);
    }
  }
}


template <typename OutT>
__device__ __forceinline__ void store4(OutT* p, float a, float b, float c, float d);
template <>
__device__ __forceinline__ void store4<float>(float* p, float a, float b, float c, float d) {
  *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d);
}
template <>
__device__ __forceinline__ void store4<bf16>(bf16* p, float a, float b, float c, float d) {
  *reinterpret_cast<uint2*>(p) = make_uint2(pack_bf16x2(a, b), pack_bf16x2(c, d));
}

// identity resize: 16 pixels (48 interleaved bytes) per thread
template <typename OutT>
__global__ void __launch_bounds__(256)
ingest_identity_vec_kernel(const uint8_t* __restrict__ src, long src_frame_stride, int h, int w,
                           const int* __restrict__ frame_index, int n_out, OutT* __restrict__ dst, int swap_rb, float divisor) {
  // byte -> float through a 256-entry table: the correctly rounded fp32 division (bit-exact with numpy's /255.0 cast to
  // float32) costs ~10 instructions per value, and the first vectorised version was issue bound, not HBM bound
  __shared__ float lut[256];
  for (int v = threadIdx.x; v < 256; v += blockDim.x) lut[v] = divisor != 1.0f ? __fdiv_rn((float)v, divisor) : (float)v;
  __syncthreads();
  const int chunks = w >> 4;                                   // 16-pixel chunks per row
  const long total = (long)n_out * h * chunks;
  const long plane = (long)h * w;
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int cx = (int)(idx % chunks);
    const long row = idx / chunks;
    const int y = (int)(row % h);
    const int f = (int)(row / h);
    const int sf = frame_index ? frame_index[f] : f;
    uint4 q[3] = {make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0)};
    if (sf >= 0) {
      const uint4* p = reinterpret_cast<const uint4*>(src + (long)sf * src_frame_stride + ((long)y * w + cx * 16) * 3);
      q[0] = __ldg(p);
      q[1] = __ldg(p + 1);
      q[2] = __ldg(p + 2);
    }
    const uint8_t* b = reinterpret_cast<const uint8_t*>(q);
    OutT* o = dst + (long)f * 3 * plane + (long)y * w + cx * 16;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int sc = swap_rb ? 2 - c : c;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        float v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = lut[b[(g * 4 + k) * 3 + sc]];
        store4<OutT>(o + c * plane + g * 4, v[0], v[1], v[2], v[3]);
      }
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
  const bool src16 = ((uintptr_t)src & 15) == 0 && (src_frame_stride & 15) == 0;
  const bool dst16 = ((uintptr_t)dst & 15) == 0;
  if (mode == 0 && (src_w & 15) == 0 && src16 && dst16) {
    const long items = (long)n_out * src_h * (src_w >> 4);
    const long cap = (long)b2_num_sms() * 16;
    const int g = (int)((items + 255) / 256 < cap ? (items + 255) / 256 : cap);
    if (out_bf16)
      ingest_identity_vec_kernel<bf16><<<g, 256, 0, st>>>((const uint8_t*)src, src_frame_stride, src_h, src_w, frame_index,
                                                         n_out, (bf16*)dst, swap_rb, divisor);
    else
      ingest_identity_vec_kernel<float><<<g, 256, 0, st>>>((const uint8_t*)src, src_frame_stride, src_h, src_w, frame_index,
                                                          n_out, (float*)dst, swap_rb, divisor);
    B2_LAUNCH_CHECK("ingest_identity_vec_kernel");
    return 0;
  }
  if (out_bf16)
    ingest_kernel<bf16><<<grid, 256, 0, st>>>((const uint8_t*)src, src_frame_stride, src_h, src_w, frame_index, n_out,
                                              (bf16*)dst, dst_h, dst_w, swap_rb, divisor, mode, sx, sy);
  else
    ingest_kernel<float><<<grid, 256, 0, st>>>((const uint8_t*)src, src_frame_stride, src_h, src_w, frame_index,
                                               n_out, (float*)dst, dst_h, dst_w, swap_rb, divisor, mode, sx, sy);
  B2_LAUNCH_CHECK("ingest_kernel");
  return 0;
}
