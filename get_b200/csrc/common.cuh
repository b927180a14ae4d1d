// Shared helpers for the get_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/get_b200.h"

namespace getb {

// ---- status / error plumbing -------------------------------------------------------------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define GETB_REQUIRE(cond, ...)                 \
  do {                                          \
    if (!(cond)) {                              \
      getb::set_error(__VA_ARGS__);             \
      return -1;                                \
    }                                           \
  } while (0)

// Call after every kernel launch: returns the cudaError_t (as int) from the enclosing C-ABI function.
#define GETB_CHECK_LAUNCH(name)                                                  \
  do {                                                                           \
    cudaError_t e__ = cudaGetLastError();                                        \
    if (e__ != cudaSuccess) {                                                    \
      getb::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));   \
      return (int)e__;                                                           \
    }                                                                            \
    getb::count_launch();                                                        \
  } while (0)

// ---- counter-based dropout ------------------------------------------------------------------
// keep(seed, i) = (hash32(i32 + seed * 0x9E3779B9) >> 8) >= floor(p * 2^24).  Multipliers are < 2^31 so
// the host-side mirror (get_b200/dropout.py) can use int64 arithmetic without overflow.
__host__ __device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16;
  x *= 0x7feb352dU;
  x ^= x >> 15;
  x *= 0x2c1b3c6dU;
  x ^= x >> 16;
  x *= 0x297a2d39U;
  x ^= x >> 15;
  return x;
}

__host__ __device__ __forceinline__ uint32_t drop_threshold(float p) { return (uint32_t)(p * 16777216.0f); }

__host__ __device__ __forceinline__ bool drop_keep(uint32_t seed, uint64_t idx, uint32_t thr) {
  uint32_t i32 = (uint32_t)idx + (uint32_t)(idx >> 32) * 0x632be5abU;
  return (hash32(i32 + seed * 0x9E3779B9U) >> 8) >= thr;
}

// ---- math -------------------------------------------------------------------------------------
__device__ __forceinline__ float sigmoidf_(float v) { return 1.0f / (1.0f + expf(-v)); }

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

__host__ __device__ __forceinline__ bool aligned16(const void* p) { return (((uintptr_t)p) & 15u) == 0; }

static inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

}  // namespace getb
