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
// Device word added to every dropout seed (misc.cu). A captured CUDA graph replays the same kernel arguments, so the
// per-step variation of the masks comes from this word, advanced by a one-thread kernel inside the graph.
const uint32_t* dropout_salt_ptr();

// TMA tensor map (CUtensorMap, 128 bytes) over {d0 inner contiguous, d1 rows `ld` elements apart[, d2 slices `ps` apart]} with
// box {b0, b1[, b2]}; bf16 (f32 = 0) or fp32 elements; sw = swizzle span in bytes (0 = none). Memoised (gemm_bp.cu).
bool make_tensor_map(void* map_out, const void* ptr, int f32, int rank, int64_t d0, int64_t d1, int64_t d2, int64_t ld, int64_t ps,
                     int b0, int b1, int b2, int sw);

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
// One 32-bit hash serves an aligned PAIR of elements (16 bits each):
//   bits(seed, w) = mix32(lo32(w) * 0x9E3779B1 + hi32(w) * 0x632BE5AB + seed * 0x85EBCA6B + 0x6A09E667),  w = idx >> 1
//   keep(seed, idx) = ((bits >> (16 * (idx & 1))) & 0xFFFF) >= thr16,   thr16 = floor(p * 65536)
// so a float4 of 4 consecutive elements (idx % 4 == 0) costs two hashes. get_b200/dropout.py mirrors it on the host.
__host__ __device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 15;
  x *= 0x2c1b3c6dU;
  x ^= x >> 12;
  x *= 0x297a2d39U;
  x ^= x >> 15;
  return x;
}

__host__ __device__ __forceinline__ uint32_t drop_threshold(float p) { return (uint32_t)(p * 65536.0f); }

__host__ __device__ __forceinline__ uint32_t drop_bits(uint32_t seed, uint64_t word) {
  return mix32((uint32_t)word * 0x9E3779B1U + (uint32_t)(word >> 32) * 0x632be5abU + seed * 0x85EBCA6BU + 0x6A09E667U);
}

__host__ __device__ __forceinline__ bool drop_keep(uint32_t seed, uint64_t idx, uint32_t thr) {
  const uint32_t b = drop_bits(seed, idx >> 1);
  return ((b >> (16u * (uint32_t)(idx & 1))) & 0xFFFFu) >= thr;
}

// 4 consecutive elements starting at idx4 (idx4 % 4 == 0): f <- keep ? f * scale : 0
__device__ __forceinline__ void drop_apply4(uint32_t seed, uint64_t idx4, uint32_t thr, float scale, float4& f) {
  const uint32_t b0 = drop_bits(seed, idx4 >> 1);
  const uint32_t b1 = drop_bits(seed, (idx4 >> 1) + 1);
  f.x = (b0 & 0xFFFFu) >= thr ? f.x * scale : 0.f;
  f.y = (b0 >> 16) >= thr ? f.y * scale : 0.f;
  f.z = (b1 & 0xFFFFu) >= thr ? f.z * scale : 0.f;
  f.w = (b1 >> 16) >= thr ? f.w * scale : 0.f;
}

// ---- math -------------------------------------------------------------------------------------
__device__ __forceinline__ float sigmoidf_(float v) { return 1.0f / (1.0f + expf(-v)); }
// Epilogue activations: ex2.approx / rcp.approx based, absolute error ~2e-7 on outputs in [-1, 1] (fp32 round-off level)
__device__ __forceinline__ float sigmoid_fast(float v) { return __fdividef(1.0f, 1.0f + __expf(-v)); }
__device__ __forceinline__ float tanh_fast(float v) {
  const float c = fminf(fmaxf(v, -15.0f), 15.0f);
  return 1.0f - __fdividef(2.0f, 1.0f + __expf(2.0f * c));
}

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
