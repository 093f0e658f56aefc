// Shared helpers for libfami_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <atomic>
#include <stdio.h>
#include <string.h>

#include "../../include/fami_b200.h"

namespace fami {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define FAMI_CHECK_ARG(cond, ...)            \
  do {                                       \
    if (!(cond)) {                           \
      fami::set_error(__VA_ARGS__);          \
      return 1;                              \
    }                                        \
  } while (0)

#define FAMI_CHECK_LAUNCH(name)                                                   \
  do {                                                                            \
    cudaError_t e__ = cudaGetLastError();                                         \
    if (e__ != cudaSuccess) {                                                     \
      fami::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));    \
      return 2;                                                                   \
    }                                                                             \
    fami::count_launch();                                                         \
  } while (0)

static inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ---- storage <-> float ---------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float to_f<__half>(__half v) { return __half2float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ __half from_f<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// load 4 consecutive channels (16B for float, 8B for bf16); pointer must be aligned accordingly
template <typename T> __device__ __forceinline__ float4 ld4(const T* p);
template <> __device__ __forceinline__ float4 ld4<float>(const float* p) {
  return __ldg(reinterpret_cast<const float4*>(p));
}
template <> __device__ __forceinline__ float4 ld4<__nv_bfloat16>(const __nv_bfloat16* p) {
  uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&u.x);
  __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&u.y);
  float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}
template <> __device__ __forceinline__ float4 ld4<__half>(const __half* p) {
  uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
  float2 fa = __half22float2(*reinterpret_cast<__half2*>(&u.x)), fb = __half22float2(*reinterpret_cast<__half2*>(&u.y));
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}
template <typename T> __device__ __forceinline__ void st4(T* p, float4 v);
template <> __device__ __forceinline__ void st4<__half>(__half* p, float4 v) {
  __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&a);
  u.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = u;
}
template <> __device__ __forceinline__ void st4<float>(float* p, float4 v) {
  *reinterpret_cast<float4*>(p) = v;
}
template <> __device__ __forceinline__ void st4<__nv_bfloat16>(__nv_bfloat16* p, float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y);
  __nv_bfloat162 b = __floats2bfloat162_rn(v.z, v.w);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&a);
  u.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = u;
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

// cp.async 16B with zero-fill when !pred (src-size 0)
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool pred) {
  uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  int sz = pred ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gsrc), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

int num_sms();   // SM count of the CURRENT device (cached per device)

// cudaFuncSetAttribute applies to the current device only: opt a kernel in to large dynamic shared memory once per
// DEVICE (not once per process), so a single process driving several GPUs (nn.DataParallel, which the reference's
// trainer uses) launches correctly on every one of them.  `mask` is a per-kernel static; racing callers at worst
// repeat the (idempotent) driver call.
template <typename K>
static inline void set_max_smem_once(std::atomic<uint64_t>& mask, K kernel, int bytes) {
  int dev = 0;
  cudaGetDevice(&dev);
  const uint64_t bit = 1ull << (dev & 63);
  if (mask.load(std::memory_order_acquire) & bit) return;
  cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  mask.fetch_or(bit, std::memory_order_release);
}

static inline bool is_half_dtype(int dt) { return dt == FAMI_BF16 || dt == FAMI_F16; }
static inline bool is_tc_dtype(int dt) { return dt == FAMI_BF16 || dt == FAMI_F16 || dt == FAMI_TF32; }
static inline size_t dtype_size(int dt) { return (dt == FAMI_F32 || dt == FAMI_TF32) ? 4 : 2; }

}  // namespace fami
