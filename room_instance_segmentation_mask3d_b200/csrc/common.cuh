// Shared helpers for libmk3d.so (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <string>

#include "../../include/mk3d.h"

namespace mk {

extern thread_local std::string g_last_error;

inline int fail(const char* where, const char* what) {
  g_last_error = std::string(where) + ": " + what;
  return 1;
}

inline int check_launch(const char* where) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(where, cudaGetErrorString(e));
  return 0;
}

#define MK_REQUIRE(cond, msg) \
  do {                        \
    if (!(cond)) return mk::fail(__func__, msg); \
  } while (0)

inline int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- coordinate key packing: (batch:10 | x:18 | y:18 | z:18), coordinates biased by 2^17 --------
constexpr int kCoordBits = 18;
constexpr int kCoordBias = 1 << 17;
constexpr unsigned long long kEmptyKey = ~0ull;

__device__ __forceinline__ bool coord_in_range(int b, int x, int y, int z) {
  const unsigned lim = 1u << kCoordBits;
  return (unsigned)b < 1023u && (unsigned)(x + kCoordBias) < lim && (unsigned)(y + kCoordBias) < lim &&
         (unsigned)(z + kCoordBias) < lim;
}

__device__ __forceinline__ unsigned long long pack_key(int b, int x, int y, int z) {
  return ((unsigned long long)(unsigned)b << (3 * kCoordBits)) |
         ((unsigned long long)(unsigned)(x + kCoordBias) << (2 * kCoordBits)) |
         ((unsigned long long)(unsigned)(y + kCoordBias) << kCoordBits) | (unsigned long long)(unsigned)(z + kCoordBias);
}

__device__ __forceinline__ unsigned long long hash_key(unsigned long long k) {
  // murmur3 finaliser
  k ^= k >> 33;
  k *= 0xff51afd7ed558ccdULL;
  k ^= k >> 33;
  k *= 0xc4ceb9fe1a85ec53ULL;
  k ^= k >> 33;
  return k;
}

// floor division for possibly negative numerators (ME floors in float; exact for |x| < 2^24)
__device__ __forceinline__ int floor_div(int a, int s) {
  int q = a / s;
  return (a % s != 0 && ((a < 0) != (s < 0))) ? q - 1 : q;
}

// ---- dtype helpers ------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) {
  return v;
}
template <>
__device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) {
  return __bfloat162float(v);
}
template <typename T>
__device__ __forceinline__ T from_f32(float v);
template <>
__device__ __forceinline__ float from_f32<float>(float v) {
  return v;
}
template <>
__device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) {
  return __float2bfloat16_rn(v);
}

// load 4 consecutive channels as fp32
__device__ __forceinline__ float4 load4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 load4(const __nv_bfloat16* p) {
  uint2 r = *reinterpret_cast<const uint2*>(p);
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&r.x);
  __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&r.y);
  float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}
__device__ __forceinline__ void store4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void store4(__nv_bfloat16* p, float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
  uint2 r;
  r.x = *reinterpret_cast<uint32_t*>(&a);
  r.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = r;
}

// 32-byte store (sm_100: STG.256): a whole sector per lane, p 32-byte aligned
__device__ __forceinline__ void stg256(void* p, const uint32_t (&w)[8]) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]),
               "r"(w[5]), "r"(w[6]), "r"(w[7])
               : "memory");
}
// 16 consecutive channels as whole sectors: 2 x 32-byte stores (fp32) or 1 (bf16); p 32-byte aligned
__device__ __forceinline__ void store16_sectors(float* p, const float (&v)[16]) {
#pragma unroll
  for (int j = 0; j < 16; j += 8) {
    uint32_t w[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) w[t] = __float_as_uint(v[j + t]);
    stg256(p + j, w);
  }
}
__device__ __forceinline__ void store16_sectors(__nv_bfloat16* p, const float (&v)[16]) {
  uint32_t w[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    __nv_bfloat162 t = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
    w[j] = *reinterpret_cast<uint32_t*>(&t);
  }
  stg256(p, w);
}

// 16 consecutive channels from fp32 registers: 4 x 16-byte stores (fp32) or 2 x 16-byte stores (bf16)
__device__ __forceinline__ void store16(float* p, const float (&v)[16]) {
#pragma unroll
  for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(p + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
}
__device__ __forceinline__ void store16(__nv_bfloat16* p, const float (&v)[16]) {
  uint32_t w[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    __nv_bfloat162 t = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
    w[j] = *reinterpret_cast<uint32_t*>(&t);
  }
  *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
  *reinterpret_cast<uint4*>(p + 8) = make_uint4(w[4], w[5], w[6], w[7]);
}

// scene of a row: largest b with offsets[b] <= row  (B is small; linear scan of a broadcast-cached array)
__device__ __forceinline__ int scene_of_row(const int32_t* __restrict__ offsets, int B, int64_t row) {
  int b = 0;
  while (b + 1 < B && row >= offsets[b + 1]) ++b;
  return b;
}

}  // namespace mk
