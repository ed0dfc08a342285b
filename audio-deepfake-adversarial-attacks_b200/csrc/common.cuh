// Shared helpers for libadvb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <string>

namespace advb {

void set_error(const std::string& msg);

struct LaunchCounter {
  int64_t n = 0;
};
extern thread_local LaunchCounter* g_counter;
inline void count_launch(int k = 1) {
  if (g_counter) g_counter->n += k;
}
// Convolution schedule of the current call (engine option "conv_sched"): 0 = persistent warp-specialised kernels
// (default), 1 = one-tile-per-CTA kernels of conv_tc.cu only (the first tcgen05 version, kept as an in-process cross-check).
extern thread_local int g_conv_sched;

// Optional per-kernel timing (bench.py's live roofline): when enabled, an event is recorded after every launch;
// consecutive events on the (serial) stream bracket exactly one kernel.
void prof_mark(const char* name, cudaStream_t stream);

#define ADVB_CUDA_OK(expr)                                                                              \
  do {                                                                                                  \
    cudaError_t _e = (expr);                                                                            \
    if (_e != cudaSuccess) {                                                                            \
      ::advb::set_error(std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " at " + __FILE__ + \
                        ":" + std::to_string(__LINE__));                                                \
      return 1;                                                                                         \
    }                                                                                                   \
  } while (0)

#define ADVB_KERNEL_OK(name, stream)                                                              \
  do {                                                                                            \
    cudaError_t _e = cudaGetLastError();                                                          \
    if (_e != cudaSuccess) {                                                                      \
      ::advb::set_error(std::string("kernel launch failed (") + (name) + "): " +                  \
                        cudaGetErrorString(_e) + " at " + __FILE__ + ":" + std::to_string(__LINE__)); \
      return 1;                                                                                   \
    }                                                                                             \
    ::advb::count_launch();                                                                       \
    ::advb::prof_mark((name), (stream));                                                          \
  } while (0)

#define ADVB_CHECK(cond, msg)                                                       \
  do {                                                                              \
    if (!(cond)) {                                                                  \
      ::advb::set_error(std::string(msg) + " (" #cond ") at " + __FILE__ + ":" +    \
                        std::to_string(__LINE__));                                  \
      return 1;                                                                     \
    }                                                                               \
  } while (0)

#define ADVB_TRY(expr)       \
  do {                       \
    int _r = (expr);         \
    if (_r != 0) return _r;  \
  } while (0)

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
static inline int64_t cdiv64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- device helpers -------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Order-preserving float <-> int key so a float max can use integer atomicMax.
__device__ __forceinline__ int float_to_key(float f) {
  int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float key_to_float(int k) {
  return __int_as_float(k >= 0 ? k : k ^ 0x7fffffff);
}

// Activation (NHWC, zero spatial border of `pad` pixels) descriptor.
struct Act {
  float* p;
  int H, W, C, pad;
  __host__ __device__ int Hp() const { return H + 2 * pad; }
  __host__ __device__ int Wp() const { return W + 2 * pad; }
  __host__ __device__ size_t per_clip() const { return (size_t)Hp() * Wp() * C; }
};

}  // namespace advb
