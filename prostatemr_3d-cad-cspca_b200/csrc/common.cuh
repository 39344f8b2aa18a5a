// common.cuh — shared helpers for the sm_100a kernels of libm1b200.so
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
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
  float* scratch;       // small fp32 scratch (per-(sample, channel) reduction results)
  size_t scratch_bytes;
  // deterministic reductions (norm_se.cu, attn_latent_loss.cu): per-block partial sums + ticket counters that
  // the last block resets to zero - one reduction kernel at a time per context (one stream per context)
  float* partial;
  size_t partial_bytes;
  unsigned int* counters;
  size_t counter_bytes;
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

// ---- typed element access (activations are fp32, bf16 or fp16) ---------------------------
// 16-bit element types share one code path: cvt2<T> unpacks / pack2<T> packs two elements of a 32-bit word.
template <typename T> __device__ __forceinline__ float2 cvt2(uint32_t w);
template <> __device__ __forceinline__ float2 cvt2<__nv_bfloat16>(uint32_t w) {
  return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xFFFF0000u));   // bf16 = high half of fp32
}
template <> __device__ __forceinline__ float2 cvt2<__half>(uint32_t w) {
  return __half22float2(*reinterpret_cast<const __half2*>(&w));
}
template <typename T> __device__ __forceinline__ uint32_t pack2(float a, float b);
template <> __device__ __forceinline__ uint32_t pack2<__nv_bfloat16>(float a, float b) {
  const __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&t);
}
template <> __device__ __forceinline__ uint32_t pack2<__half>(float a, float b) {
  const __half2 t = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&t);
}

template <typename T> __device__ __forceinline__ float ld_f(const T* p);
template <> __device__ __forceinline__ float ld_f<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ld_f<__nv_bfloat16>(const __nv_bfloat16* p) {
  return __bfloat162float(*p);
}
template <> __device__ __forceinline__ float ld_f<__half>(const __half* p) { return __half2float(*p); }
template <typename T> __device__ __forceinline__ void st_f(T* p, float v);
template <> __device__ __forceinline__ void st_f<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void st_f<__nv_bfloat16>(__nv_bfloat16* p, float v) {
  *p = __float2bfloat16_rn(v);
}
template <> __device__ __forceinline__ void st_f<__half>(__half* p, float v) { *p = __float2half_rn(v); }

// 8-wide vector access: fp32 -> 2 x float4, 16-bit types -> one 16-byte access
template <typename T> __device__ __forceinline__ void ld8v(const T* p, float (&v)[8]) {
  if constexpr (sizeof(T) == 4) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float2 f = cvt2<T>(w[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
  }
}
template <typename T> __device__ __forceinline__ void st8v(T* p, const float (&v)[8]) {
  if constexpr (sizeof(T) == 4) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
  } else {
    *reinterpret_cast<uint4*>(p) = make_uint4(pack2<T>(v[0], v[1]), pack2<T>(v[2], v[3]), pack2<T>(v[4], v[5]),
                                              pack2<T>(v[6], v[7]));
  }
}

// 4-wide vector access: fp32 -> float4 (16 B), 16-bit types -> 8 B
template <typename T> struct Vec4 {
  static __device__ __forceinline__ void load(const T* p, float (&v)[4]) {
    if constexpr (sizeof(T) == 4) {
      const float4 t = *reinterpret_cast<const float4*>(p);
      v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else {
      const uint2 t = *reinterpret_cast<const uint2*>(p);
      const float2 a = cvt2<T>(t.x), b = cvt2<T>(t.y);
      v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
    }
  }
  static __device__ __forceinline__ void store(T* p, const float (&v)[4]) {
    if constexpr (sizeof(T) == 4) {
      *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
      *reinterpret_cast<uint2*>(p) = make_uint2(pack2<T>(v[0], v[1]), pack2<T>(v[2], v[3]));
    }
  }
};

// ---- dtype dispatch ---------------------------------------------------------------------
// M1_DISPATCH_T(code, T, ...): runs the statement with `using T` = float / __nv_bfloat16 / __half.
// M1_DISPATCH_VG(code, TV, TG, ...): value type TV of `code` and the type TG its GRADIENTS are stored in
// (m1_grad_dtype: fp32 -> fp32, bf16 -> bf16, fp16 -> bf16 - fp16 has too little range for gradients).
#define M1_DISPATCH_T(code, T, ...)                                                   \
  do {                                                                                \
    if ((code) == M1_BF16) { using T = __nv_bfloat16; __VA_ARGS__; }                  \
    else if ((code) == M1_F16) { using T = __half; __VA_ARGS__; }                     \
    else { using T = float; __VA_ARGS__; }                                            \
  } while (0)
#define M1_DISPATCH_VG(code, TV, TG, ...)                                             \
  do {                                                                                \
    if ((code) == M1_BF16) { using TV = __nv_bfloat16; using TG = __nv_bfloat16; __VA_ARGS__; } \
    else if ((code) == M1_F16) { using TV = __half; using TG = __nv_bfloat16; __VA_ARGS__; }    \
    else { using TV = float; using TG = float; __VA_ARGS__; }                         \
  } while (0)
#define M1_DISPATCH_T2(code_a, code_b, TA, TB_, ...) \
  M1_DISPATCH_T(code_a, TA, M1_DISPATCH_T(code_b, TB_, __VA_ARGS__))
static inline bool m1_is16(int code) { return code == M1_BF16 || code == M1_F16; }
static inline int m1_grad_dtype_of(int code) { return code == M1_F16 ? M1_BF16 : code; }
// element type of the packed tensor-core weight operand of a convolution launch
static inline int m1_conv_w_dtype(const m1_conv_desc* d) { return d->w_dtype ? d->w_dtype : d->act_dtype; }

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

// Deterministic grid-wide sum of one value per block, added to *out (scaled) by the LAST block to finish:
// block partials go to `part[blockIdx.x]`, a self-resetting ticket counter elects the last block, which sums the
// partials in a fixed order. `v` must hold the block's value in thread 0. All threads of the block must call.
struct GridSum {
  float* part;            // >= gridDim.x floats (m1_ctx::partial)
  unsigned int* counter;  // zero between launches (m1_ctx::counters + kGridSumCounter)
};
constexpr int kGridSumCounter = 8192;     // counters [0, 8192) belong to the per-sample reductions
template <int BLOCK>
__device__ __forceinline__ void grid_sum_add(float v, float scale, float* out, GridSum gs, float* smem /* BLOCK/32 */) {
  __shared__ unsigned int s_ticket_gs;
  if (threadIdx.x == 0) {
    gs.part[blockIdx.x] = v;
    __threadfence();
    s_ticket_gs = atomicInc(gs.counter, gridDim.x - 1);
  }
  __syncthreads();
  if (s_ticket_gs != gridDim.x - 1) return;
  __threadfence();
  float t[1] = {0.f};
  for (unsigned i = threadIdx.x; i < gridDim.x; i += BLOCK) t[0] += __ldcg(gs.part + i);
  block_sum<1, BLOCK>(t, smem);
  if (threadIdx.x == 0) *out += t[0] * scale;
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
static inline GridSum m1_grid_sum(const m1_ctx* ctx) {
  return GridSum{ctx->partial, ctx->counters + kGridSumCounter};
}

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
