// Shared helpers for the b200lrcn kernels (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#define B2_API extern "C" __attribute__((visibility("default")))

// ---- status plumbing (include/b200lrcn.h: 0 ok, <0 argument error, >0 cudaError_t) ----
void b2_set_error(const char* fmt, ...);
void b2_count_launch(int n = 1);

#define B2_ARG_CHECK(cond, ...)                 \
  do {                                          \
    if (!(cond)) {                              \
      b2_set_error(__VA_ARGS__);                \
      return -1;                                \
    }                                           \
  } while (0)

#define B2_LAUNCH_CHECK(name)                                                    \
  do {                                                                           \
    cudaError_t e__ = cudaGetLastError();                                        \
    if (e__ != cudaSuccess) {                                                    \
      b2_set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));      \
      return (int)e__;                                                           \
    }                                                                            \
    b2_count_launch();                                                           \
  } while (0)

#define B2_CUDA_CHECK(call)                                                      \
  do {                                                                           \
    cudaError_t e__ = (call);                                                    \
    if (e__ != cudaSuccess) {                                                    \
      b2_set_error("%s failed: %s", #call, cudaGetErrorString(e__));             \
      return (int)e__;                                                           \
    }                                                                            \
  } while (0)

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is PER DEVICE: a process that touches a second GPU must repeat it
// there.  One flag word per call site, one bit (or one high-water mark) per device ordinal; safe from concurrent host threads.
constexpr int kB2MaxDevices = 64;
static inline int b2_current_device() {
  int d = 0;
  cudaGetDevice(&d);
  return d & (kB2MaxDevices - 1);
}
struct B2PerDeviceOnce {
  std::atomic<unsigned long long> done{0};
  bool needed() const { return !((done.load(std::memory_order_acquire) >> b2_current_device()) & 1ull); }
  void mark() { done.fetch_or(1ull << b2_current_device(), std::memory_order_release); }
};
struct B2PerDeviceMax {      // largest dynamic shared-memory size the attribute was raised to, per device
  std::atomic<int> v[kB2MaxDevices] = {};
  bool below(int need) const { return v[b2_current_device()].load(std::memory_order_acquire) < need; }
  void set(int need) { v[b2_current_device()].store(need, std::memory_order_release); }
};

static inline int b2_ceil_div(long a, long b) { return (int)((a + b - 1) / b); }
int b2_num_sms();

typedef __nv_bfloat16 bf16;

// ---- small device helpers ----
__device__ __forceinline__ float bf16_round(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&p);
}
__device__ __forceinline__ void unpack_bf16x2(uint32_t u, float& lo, float& hi) {
  __nv_bfloat162 p = *reinterpret_cast<__nv_bfloat162*>(&u);
  float2 f = __bfloat1622float2(p);
  lo = f.x;
  hi = f.y;
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }
// tanh(x) = 2 sigmoid(2x) - 1 on the fast exponential: ~1e-7 absolute error, a short dependent chain (the
// library tanhf is ~30 dependent instructions and sits twice on the LSTM recurrence's critical path)
__device__ __forceinline__ float tanhf_(float x) { return 2.0f / (1.0f + __expf(-2.0f * x)) - 1.0f; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
