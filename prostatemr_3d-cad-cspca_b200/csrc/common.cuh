// common.cuh — shared helpers for the sm_100a kernels of libm1b200.so
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string.h>
#include "../../include/m1b200.h"

#define M1_LRELU_SLOPE 0.1f

struct m1_ctx {
  int device;
  int num_sms;
  int64_t launches;
  void* encode_tiled;   // cuTensorMapEncodeTiled entry point (driver API, fetched at runtime)
  float* scratch;       // small fp32 scratch (reductions)
  size_t scratch_bytes;
};

void m1_set_error(const char* fmt, ...);

#define M1_CHECK(cond, ...)                          \
  do {                                               \
    if (!(cond)) {                                   \
      m1_set_error(__VA_ARGS__);                     \
      return 1;                                      \
    }                                                \
  } while (0)

#define M1_CUDA(expr)                                                                   \
  do {                                                                                  \
    cudaError_t e__ = (expr);                                                           \
    if (e__ != cudaSuccess) {                                                           \
      m1_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__,   \
                   __LINE__);                                                           \
      return 1;                                                                         \
    }                                                                                   \
  } while (0)

#define M1_LAUNCH_CHECK(ctx)                                                            \
  do {                                                                                  \
    cudaError_t e__ = cudaGetLastError();                                               \
    if (e__ != cudaSuccess) {                                                           \
      m1_set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__),         \
                   __FILE__, __LINE__);                                                 \
      return 1;                                                                         \
    }                                                                                   \
    (ctx)->launches++;                                                                  \
  } while (0)

// ---- typed element access (activations are fp32 or bf16) -------------------------------
template <typename T> __device__ __forceinline__ float ld_f(const T* p);
template <> __device__ __forceinline__ float ld_f<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ld_f<__nv_bfloat16>(const __nv_bfloat16* p) {
  return __bfloat162float(*p);
}
template <typename T> __device__ __forceinline__ void st_f(T* p, float v);
template <> __device__ __forceinline__ void st_f<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void st_f<__nv_bfloat16>(__nv_bfloat16* p, float v) {
  *p = __float2bfloat16_rn(v);
}

// 4-wide vector access: fp32 -> float4 (16 B), bf16 -> 8 B
template <typename T> struct Vec4;
template <> struct Vec4<float> {
  static __device__ __forceinline__ void load(const float* p, float (&v)[4]) {
    float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
template <> struct Vec4<__nv_bfloat16> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[4]) {
    uint2 t = *reinterpret_cast<const uint2*>(p);
    __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&t.x);
    __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&t.y);
    v[0] = __low2float(a); v[1] = __high2float(a); v[2] = __low2float(b); v[3] = __high2float(b);
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[4]) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]);
    __nv_bfloat162 b = __floats2bfloat162_rn(v[2], v[3]);
    uint2 t;
    t.x = *reinterpret_cast<uint32_t*>(&a);
    t.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = t;
  }
};

__device__ __forceinline__ float lrelu(float x, float slope) { return x > 0.f ? x : x * slope; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum of up to NV values per thread; result valid in thread 0 (and warp 0)
template <int NV, int BLOCK>
__device__ __forceinline__ void block_sum(float (&v)[NV], float* smem /* NV * BLOCK/32 */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) smem[i * (BLOCK / 32) + warp] = v[i];
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float t = lane < BLOCK / 32 ? smem[i * (BLOCK / 32) + lane] : 0.f;
      v[i] = warp_sum(t);
    }
  }
  __syncthreads();
}

// ---- Philox4x32-10 counter-based RNG (dropout masks regenerated in backward) -------------
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
  uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
  uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}
// the 4 raw 32-bit words of Philox4x32-10 for counter (idx4, stream_id) under key seed
__device__ __forceinline__ void philox_words4(uint64_t seed, uint64_t stream_id, uint64_t idx4, uint32_t (&w)[4]) {
  uint32_t c[4] = {(uint32_t)idx4, (uint32_t)(idx4 >> 32), (uint32_t)stream_id, (uint32_t)(stream_id >> 32)};
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) w[i] = c[i];
}
// 4 uniforms in [0,1) for counter (idx4, stream_id) under key seed
__device__ __forceinline__ void philox_uniform4(uint64_t seed, uint64_t stream_id, uint64_t idx4,
                                                float (&u)[4]) {
  uint32_t c[4] = {(uint32_t)idx4, (uint32_t)(idx4 >> 32), (uint32_t)stream_id,
                   (uint32_t)(stream_id >> 32)};
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) u[i] = (float)(c[i] >> 8) * (1.0f / 16777216.0f);
}

static inline int64_t cdiv64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// entry points implemented per translation unit
int m1_conv3d_simt(m1_ctx* ctx, const m1_conv_desc* d, const void* const* srcs,
                   const float* const* w, const float* const* bias, void* const* outs,
                   cudaStream_t st);
int m1_conv3d_tc(m1_ctx* ctx, const m1_conv_desc* d, const void* const* srcs,
                 const void* w_packed, const float* const* bias, void* const* outs,
                 cudaStream_t st);
int m1_conv3d_halo_supported(const m1_conv_desc* d, int* preferred);
int m1_conv3d_tc_plan_info(const m1_conv_desc* d, int32_t* out);
int m1_conv3d_halo_plan_info(const m1_conv_desc* d, int32_t* out);
int m1_conv3d_wgrad_plan_info(const m1_conv_desc* d, int32_t* out);
int m1_conv3d_halo(m1_ctx* ctx, const m1_conv_desc* d, const void* const* srcs, const void* w_packed,
                   const float* const* bias, void* const* outs, cudaStream_t st);
int m1_conv3d_wgrad_tc_supported(const m1_conv_desc* d, int j0, int jn);
int m1_conv3d_wgrad_tc(m1_ctx* ctx, const m1_conv_desc* d, int j0, int jn, const void* const* srcs,
                       const void* const* douts, float* const* dws, cudaStream_t st);
