// optim_util.cu — K9 fused Adam-AMSGrad (+ L2 regulariser) and the small host utilities.
//   K9  tf.keras.optimizers.Adam(amsgrad=True) dense update (train_model.py:113-120) with the
//       Keras l2 regularisers of R:networks.py:259-263 folded in as g += 2*l2*w.
#include "common.cuh"
#include <algorithm>

namespace {
constexpr int TB = 256;

__global__ void __launch_bounds__(TB) adam_kernel(float* __restrict__ w, const float* __restrict__ g,
                                                 float* __restrict__ m, float* __restrict__ v,
                                                 float* __restrict__ vhat, int64_t n, float lr_t,
                                                 const float* __restrict__ lr_dev, float b1,
                                                 float b2, float eps, float l2, float gscale, int amsgrad,
                                                 float* __restrict__ l2_out, GridSum gs) {
  __shared__ float sm[TB / 32];
  if (lr_dev != nullptr) lr_t = *lr_dev;      // graph replay: the step size lives in device memory
  float acc[1] = {0.f};
  const int64_t n4 = n / 4;
  for (int64_t i = blockIdx.x * (int64_t)TB + threadIdx.x; i < n4; i += (int64_t)gridDim.x * TB) {
    float4 W = reinterpret_cast<float4*>(w)[i];
    const float4 G = reinterpret_cast<const float4*>(g)[i];
    float4 M = reinterpret_cast<float4*>(m)[i];
    float4 V = reinterpret_cast<float4*>(v)[i];
    float4 H = reinterpret_cast<float4*>(vhat)[i];
    float* pw = &W.x; const float* pg = &G.x; float* pm = &M.x; float* pv = &V.x; float* ph = &H.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      acc[0] = fmaf(pw[k], pw[k], acc[0]);
      const float gg = fmaf(2.f * l2, pw[k], pg[k] * gscale);
      pm[k] = b1 * pm[k] + (1.f - b1) * gg;
      pv[k] = b2 * pv[k] + (1.f - b2) * gg * gg;
      ph[k] = amsgrad ? fmaxf(ph[k], pv[k]) : pv[k];      // plain Adam: the denominator is sqrt(v)
      pw[k] -= lr_t * pm[k] / (sqrtf(ph[k]) + eps);
    }
    reinterpret_cast<float4*>(w)[i] = W;
    reinterpret_cast<float4*>(m)[i] = M;
    reinterpret_cast<float4*>(v)[i] = V;
    reinterpret_cast<float4*>(vhat)[i] = H;
  }
  // tail
  for (int64_t i = n4 * 4 + blockIdx.x * (int64_t)TB + threadIdx.x; i < n; i += (int64_t)gridDim.x * TB) {
    acc[0] = fmaf(w[i], w[i], acc[0]);
    const float gg = fmaf(2.f * l2, w[i], g[i] * gscale);
    m[i] = b1 * m[i] + (1.f - b1) * gg;
    v[i] = b2 * v[i] + (1.f - b2) * gg * gg;
    vhat[i] = amsgrad ? fmaxf(vhat[i], v[i]) : v[i];
    w[i] -= lr_t * m[i] / (sqrtf(vhat[i]) + eps);
  }
  if (l2_out) {
    block_sum<1, TB>(acc, sm);
    grid_sum_add<TB>(acc[0], l2, l2_out, gs, sm);
  }
}

template <typename S, typename D>
__global__ void cast_kernel(const S* __restrict__ s, D* __restrict__ d, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    st_f<D>(d + i, ld_f<S>(s + i));
}

template <typename S, typename D>
__global__ void copy_channels_kernel(const S* __restrict__ s, int sc, int so, D* __restrict__ d, int dc, int dofs,
                                     int c, int64_t rows) {
  const int64_t total = rows * c;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / c;
    const int k = (int)(i % c);
    st_f<D>(d + r * dc + dofs + k, ld_f<S>(s + r * sc + so + k));
  }
}

template <typename T>
__global__ void axpy_kernel(const T* __restrict__ x, float a, T* __restrict__ y, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    st_f<T>(y + i, fmaf(a, ld_f<T>(x + i), ld_f<T>(y + i)));
}

__global__ void fusion_kernel(const float* __restrict__ prior, const float* __restrict__ follow, int strategy,
                              int64_t rows, float* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < rows; i += (int64_t)gridDim.x * blockDim.x) {
    const float a = prior ? prior[i] : 0.f, b = follow[i];
    float j;
    if (strategy == 0) j = b;
    else if (strategy == 1) j = 1.f - (1.f - a) * (1.f - b);
    else j = (a * b + 1e-9f) / (a * b + 1e-9f + (1.f - a) * (1.f - b));
    out[2 * i] = 1.f - j;
    out[2 * i + 1] = j;
  }
}

__global__ void philox_normal_kernel(uint64_t seed, uint64_t stream_id, const uint64_t* __restrict__ step,
                                     float* __restrict__ out, int64_t n) {
  const int64_t n4 = (n + 3) / 4;
  if (step != nullptr) stream_id += *step * M1_PHILOX_STEP_STRIDE;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float u[4];
    philox_uniform4(seed, stream_id, (uint64_t)i, u);
    float z[4];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const float r = sqrtf(-2.f * logf(fmaxf(u[2 * k], 5.9604645e-8f)));
      float sn, cs;
      sincosf(6.28318530718f * u[2 * k + 1], &sn, &cs);
      z[2 * k] = r * cs;
      z[2 * k + 1] = r * sn;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (i * 4 + k < n) out[i * 4 + k] = z[k];
  }
}

inline unsigned nb(const m1_ctx* ctx, int64_t n) {
  return (unsigned)std::max<int64_t>(1, std::min<int64_t>(cdiv64(n, TB), (int64_t)ctx->num_sms * 16));
}
}  // namespace

extern "C" int m1_adam_amsgrad(m1_ctx* ctx, float* w, const float* g, float* m, float* v, float* vhat, int64_t n,
                               float lr_t, float beta1, float beta2, float eps, float l2, float gscale,
                               float* l2_sq_out, int amsgrad, void* stream) {
  M1_CHECK((((uintptr_t)w | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v | (uintptr_t)vhat) & 15) == 0,
           "m1_adam_amsgrad: buffers must be 16-byte aligned");
  adam_kernel<<<nb(ctx, n / 4 + 1), TB, 0, (cudaStream_t)stream>>>(w, g, m, v, vhat, n, lr_t, nullptr, beta1, beta2,
                                                                   eps, l2, gscale, amsgrad, l2_sq_out, m1_grid_sum(ctx));
  M1_LAUNCH_CHECK(ctx);
  return 0;
}

extern "C" int m1_adam_amsgrad_dev(m1_ctx* ctx, float* w, const float* g, float* m, float* v, float* vhat, int64_t n,
                                   const float* lr_t_dev, float beta1, float beta2, float eps, float l2,
                                   float gscale, float* l2_sq_out, int amsgrad, void* stream) {
  M1_CHECK((((uintptr_t)w | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v | (uintptr_t)vhat) & 15) == 0,
           "m1_adam_amsgrad_dev: buffers must be 16-byte aligned");
  M1_CHECK(lr_t_dev != nullptr, "m1_adam_amsgrad_dev: lr_t_dev is NULL");
  adam_kernel<<<nb(ctx, n / 4 + 1), TB, 0, (cudaStream_t)stream>>>(w, g, m, v, vhat, n, 0.f, lr_t_dev, beta1, beta2,
                                                                   eps, l2, gscale, amsgrad, l2_sq_out, m1_grid_sum(ctx));
  M1_LAUNCH_CHECK(ctx);
  return 0;
}

extern "C" int m1_cast(m1_ctx* ctx, const void* src, int sdtype, void* dst, int ddtype, int64_t n, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  M1_DISPATCH_T2(sdtype, ddtype, S, D, (cast_kernel<S, D><<<nb(ctx, n), TB, 0, st>>>((const S*)src, (D*)dst, n)));
  M1_LAUNCH_CHECK(ctx);
  return 0;
}

extern "C" int m1_copy_channels(m1_ctx* ctx, const void* src, int sdtype, int src_c, int src_off, void* dst,
                                int ddtype, int dst_c, int dst_off, int c, int64_t rows, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n = rows * c;
  M1_CHECK(src_off + c <= src_c && dst_off + c <= dst_c, "m1_copy_channels: slice out of range");
  M1_DISPATCH_T2(sdtype, ddtype, S, D, (copy_channels_kernel<S, D><<<nb(ctx, n), TB, 0, st>>>(
                                           (const S*)src, src_c, src_off, (D*)dst, dst_c, dst_off, c, rows)));
  M1_LAUNCH_CHECK(ctx);
  return 0;
}

extern "C" int m1_axpy(m1_ctx* ctx, const void* x, int dtype, float a, void* y, int64_t n, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  M1_DISPATCH_T(dtype, T, (axpy_kernel<T><<<nb(ctx, n), TB, 0, st>>>((const T*)x, a, (T*)y, n)));
  M1_LAUNCH_CHECK(ctx);
  return 0;
}

extern "C" int m1_decision_fusion(m1_ctx* ctx, const float* prior, const float* follow, int strategy, int64_t rows,
                                  float* out, void* stream) {
  M1_CHECK(strategy >= 0 && strategy <= 2, "m1_decision_fusion: unknown strategy %d", strategy);
  fusion_kernel<<<nb(ctx, rows), TB, 0, (cudaStream_t)stream>>>(prior, follow, strategy, rows, out);
  M1_LAUNCH_CHECK(ctx);
  return 0;
}

extern "C" int m1_philox_normal(m1_ctx* ctx, uint64_t seed, uint64_t stream_id, float* out, int64_t n,
                                void* stream) {
  philox_normal_kernel<<<nb(ctx, (n + 3) / 4), TB, 0, (cudaStream_t)stream>>>(seed, stream_id, nullptr, out, n);
  M1_LAUNCH_CHECK(ctx);
  return 0;
}

extern "C" int m1_philox_normal_step(m1_ctx* ctx, uint64_t seed, uint64_t stream_id, const uint64_t* step_dev,
                                     float* out, int64_t n, void* stream) {
  philox_normal_kernel<<<nb(ctx, (n + 3) / 4), TB, 0, (cudaStream_t)stream>>>(seed, stream_id, step_dev, out, n);
  M1_LAUNCH_CHECK(ctx);
  return 0;
}
