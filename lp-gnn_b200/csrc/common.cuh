// Shared helpers for liblpgnn (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/lpgnn.h"

namespace lpgnn {

// thread-local error string behind lpgnn_last_error()
void set_error(const char* fmt, ...);
int check_device();  // LPGNN_OK or LPGNN_ENODEVICE
int sm_count();
void count_launches(int n);  // kernels enqueued by the library (lpgnn_launch_count)

#define LPGNN_CUDA_OK(expr)                                                              \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      ::lpgnn::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return LPGNN_ECUDA;                                                                \
    }                                                                                    \
  } while (0)

#define LPGNN_LAUNCH_OK()                                                                \
  do {                                                                                   \
    cudaError_t _e = cudaGetLastError();                                                 \
    if (_e != cudaSuccess) {                                                             \
      ::lpgnn::set_error("%s:%d kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
      return LPGNN_ECUDA;                                                                \
    }                                                                                    \
  } while (0)

#define LPGNN_REQUIRE(cond, ...)                                                         \
  do {                                                                                   \
    if (!(cond)) {                                                                       \
      ::lpgnn::set_error(__VA_ARGS__);                                                   \
      return LPGNN_EINVAL;                                                               \
    }                                                                                    \
  } while (0)

// Optional extras of the tensor-core transform's epilogue (bf16 output only): out = keep ? epi(acc + bias) * out_scale : 0
// with keep = (mask_act > 0) [backward of relu / dropout] and/or the inverted-dropout hash of (seed, element index).
struct EpiX {
  const void* mask_act = nullptr;   // bf16 [M,N], same layout as out
  float out_scale = 1.f;
  uint32_t drop_threshold = 0;      // P(drop) = threshold / 2^32
  uint64_t drop_seed = 0;
};

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t hash_u64(uint64_t x) {  // splitmix64 finaliser, top 32 bits
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  x ^= x >> 31;
  return (uint32_t)(x >> 32);
}
// Keep-decisions of lpgnn_dropout (shared by the stand-alone kernel and the fused GEMM epilogue).  One 64-bit hash
// serves FOUR consecutive elements (a 16-bit field each, compared against the top 16 bits of the threshold), so an
// epilogue thread that stores 8 elements pays two hashes instead of eight: P(drop) = (threshold >> 16) / 2^16.
__device__ __forceinline__ uint64_t hash64(uint64_t x) {  // splitmix64 finaliser
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
// bit k of the result = element 4*quad + k is kept
__device__ __forceinline__ uint32_t dropout_keep4(uint64_t seed, uint64_t quad, uint32_t threshold) {
  const uint64_t h = hash64(seed ^ quad * 0xD6E8FEB86659FD93ull);
  const uint32_t t = threshold >> 16, lo = (uint32_t)h, hi = (uint32_t)(h >> 32);
  return ((lo & 0xffffu) >= t ? 1u : 0u) | ((lo >> 16) >= t ? 2u : 0u) | ((hi & 0xffffu) >= t ? 4u : 0u) | ((hi >> 16) >= t ? 8u : 0u);
}
__device__ __forceinline__ bool dropout_keep(uint64_t seed, uint64_t idx, uint32_t threshold) {
  return (dropout_keep4(seed, idx >> 2, threshold) >> (idx & 3)) & 1u;
}
#endif

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 p = __floats2bfloat162_rn(a, b);  // .x = a (low half), .y = b
  return *reinterpret_cast<uint32_t*>(&p);
}

// IEEE half (LPGNN_F16: the reference's `--fp16 1` storage, utils.py:909-915 / val.py:269) -- inference only
__device__ __forceinline__ float2 f16x2_unpack(uint32_t w) { return __half22float2(*reinterpret_cast<const __half2*>(&w)); }
__device__ __forceinline__ uint32_t pack_f16(float a, float b) {
  __half2 p = __floats2half2_rn(a, b);  // .x = a (low half), .y = b
  return *reinterpret_cast<uint32_t*>(&p);
}
// Two 16-bit floats in a 32-bit word <-> two fp32, by storage type
template <typename T> struct Half16;
template <> struct Half16<__nv_bfloat16> {
  __device__ static __forceinline__ float2 unpack(uint32_t w) { return make_float2(bf16_lo(w), bf16_hi(w)); }
  __device__ static __forceinline__ uint32_t pack(float a, float b) { return pack_bf16(a, b); }
};
template <> struct Half16<__half> {
  __device__ static __forceinline__ float2 unpack(uint32_t w) { return f16x2_unpack(w); }
  __device__ static __forceinline__ uint32_t pack(float a, float b) { return pack_f16(a, b); }
};
static inline bool is_16bit(int dtype) { return dtype == LPGNN_BF16 || dtype == LPGNN_F16; }
static inline bool dtype_ok(int dtype) { return dtype == LPGNN_F32 || is_16bit(dtype); }

// 128-bit read-only streaming load / store
__device__ __forceinline__ uint4 ldg128(const void* p) {
  return __ldg(reinterpret_cast<const uint4*>(p));
}

}  // namespace lpgnn
