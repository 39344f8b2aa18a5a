// norm_se.cu — K4/K5: the bandwidth-bound kernels around every convolution of an SE-ResNet block.
//   K4  tfa InstanceNormalization (eps 1e-3, biased variance, affine) + LeakyReLU(0.1)
//       R:networks.py:473,576; R:network_blocks.py:38-44,55,58,104
//   K5  squeeze (GAP), excite (conv6 -> lrelu -> conv7 -> sigmoid), gate * residual -> lrelu
//       -> dropout, R:network_blocks.py:68-78 + R:network_blocks.py:137-143 (tf.nn.dropout)
// All tensors are [batch][voxels][C] (NDHWC flattened); statistics/parameters are fp32.
// Vectorised 8 channels per thread (16-byte bf16 / 2 x 16-byte fp32 accesses), per-(sample,channel)
// reductions go warp-shuffle-free through shared-memory accumulators and one global atomic per
// (block, channel).
#include "common.cuh"
#include <algorithm>

namespace {

constexpr int TB = 256;

// ---- generic loaders: VW = 8 / 4 channels per thread (16-byte bf16 / 2x16-byte fp32 accesses) or 1
// (scalar fallback for odd channel counts); value arrays always have 8 slots --------------------
template <typename T, int VW>
__device__ __forceinline__ void ldv(const T* p, float (&v)[8]) {
  if constexpr (VW == 8) {
    if constexpr (sizeof(T) == 2) {
      const uint4 u = *reinterpret_cast<const uint4*>(p);
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
      for (int i = 0; i < 4; ++i) { v[2 * i] = __low2float(h[i]); v[2 * i + 1] = __high2float(h[i]); }
    } else {
      const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
  } else if constexpr (VW == 4) {
    float t[4];
    Vec4<T>::load(p, t);
    v[0] = t[0]; v[1] = t[1]; v[2] = t[2]; v[3] = t[3];
  } else {
    v[0] = ld_f<T>(p);
  }
}
template <typename T, int VW>
__device__ __forceinline__ void stv(T* p, const float (&v)[8]) {
  if constexpr (VW == 8) {
    if constexpr (sizeof(T) == 2) {
      uint4 u;
      __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
      for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
      *reinterpret_cast<uint4*>(p) = u;
    } else {
      *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
    }
  } else if constexpr (VW == 4) {
    const float t[4] = {v[0], v[1], v[2], v[3]};
    Vec4<T>::store(p, t);
  } else {
    st_f<T>(p, v[0]);
  }
}
template <int VW>
__device__ __forceinline__ void ldp(const float* p, float (&v)[8]) {  // parameter vectors
  if constexpr (VW >= 4) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    if constexpr (VW == 8) {
      const float4 u = *reinterpret_cast<const float4*>(p + 4);
      v[4] = u.x; v[5] = u.y; v[6] = u.z; v[7] = u.w;
    }
  } else {
    v[0] = *p;
  }
}
// stats are [n][C][2] interleaved (mean, rstd)
template <int VW>
__device__ __forceinline__ void ld_stats(const float* st, float (&mean)[8], float (&rstd)[8]) {
#pragma unroll
  for (int i = 0; i < VW; ++i) { mean[i] = st[2 * i]; rstd[i] = st[2 * i + 1]; }
}

// Row-parallel per-(sample, channel) reduction skeleton.
// grid = (slabs, batch); thread -> channel group cg (VW channels) and row lane; each thread walks
// rows [slab*rows_per_slab, ...) with stride `lanes`.  F(row_offset, cbase, acc[K][8]) adds.
template <int K, int VW, typename F>
__device__ __forceinline__ void reduce_rows(int64_t voxels, int C, int64_t rows_per_slab, float* smem,
                                            float* gout /* [C][K] of this sample */, float scale, F f) {
  const int CG = C / VW;
  const int cgs = min(CG, TB);
  const int lanes = TB / cgs;
  const int my_lane = threadIdx.x / cgs;
  for (int i = threadIdx.x; i < K * C; i += TB) smem[i] = 0.f;
  __syncthreads();
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_slab;
  const int64_t r1 = min(r0 + rows_per_slab, voxels);
  if (my_lane < lanes) {
    for (int cg = threadIdx.x % cgs; cg < CG; cg += cgs) {
      float acc[K][8];
#pragma unroll
      for (int k = 0; k < K; ++k)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[k][i] = 0.f;
      const int cbase = cg * VW;
      for (int64_t r = r0 + my_lane; r < r1; r += lanes) f(r, cbase, acc);
#pragma unroll
      for (int k = 0; k < K; ++k)
#pragma unroll
        for (int i = 0; i < VW; ++i) atomicAdd(&smem[k * C + cbase + i], acc[k][i]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < K * C; i += TB) {
    const int k = i / C, c = i % C;
    atomicAdd(&gout[c * K + k], smem[i] * scale);
  }
}

// ---------------------------------------------------------------------------------------------
// K4 forward
// ---------------------------------------------------------------------------------------------
template <typename T, int VW>
__global__ void __launch_bounds__(TB) inorm_sums_kernel(const T* __restrict__ x, int64_t voxels, int C,
                                                       int64_t rows_per_slab, float* __restrict__ sums) {
  extern __shared__ float smem[];
  const int n = blockIdx.y;
  const T* xb = x + (int64_t)n * voxels * C;
  reduce_rows<2, VW>(voxels, C, rows_per_slab, smem, sums + (int64_t)n * C * 2, 1.f,
                     [&](int64_t r, int cbase, float (&acc)[2][8]) {
                       float v[8];
                       ldv<T, VW>(xb + r * C + cbase, v);
#pragma unroll
                       for (int i = 0; i < VW; ++i) { acc[0][i] += v[i]; acc[1][i] = fmaf(v[i], v[i], acc[1][i]); }
                     });
}

__global__ void inorm_finalize_kernel(float* __restrict__ stats, int total, float inv_v, float eps) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const float s1 = stats[2 * i], s2 = stats[2 * i + 1];
  const float mean = s1 * inv_v;
  const float var = fmaxf(s2 * inv_v - mean * mean, 0.f);
  stats[2 * i] = mean;
  stats[2 * i + 1] = rsqrtf(var + eps);
}

template <typename T, int VW>
__global__ void __launch_bounds__(TB) inorm_act_fwd_kernel(const T* __restrict__ x, const float* __restrict__ stats,
                                                          const float* __restrict__ gamma,
                                                          const float* __restrict__ beta, int64_t voxels, int C,
                                                          float slope, T* __restrict__ y, int64_t total_vec) {
  const int CG = C / VW;
  for (int64_t i = blockIdx.x * (int64_t)TB + threadIdx.x; i < total_vec; i += (int64_t)gridDim.x * TB) {
    const int cbase = (int)(i % CG) * VW;
    const int n = (int)(i / ((int64_t)CG * voxels));
    float v[8], mean[8], rstd[8], g[8], b[8];
    ldv<T, VW>(x + i * VW, v);
    ld_stats<VW>(stats + ((int64_t)n * C + cbase) * 2, mean, rstd);
    ldp<VW>(gamma + cbase, g);
    ldp<VW>(beta + cbase, b);
#pragma unroll
    for (int k = 0; k < VW; ++k) v[k] = lrelu((v[k] - mean[k]) * rstd[k] * g[k] + b[k], slope);
    stv<T, VW>(y + i * VW, v);
  }
}

// ---------------------------------------------------------------------------------------------
// K4 backward: red[n][c] = (sum g, sum g*xhat), g = dy * act'(y)
// ---------------------------------------------------------------------------------------------
template <typename T, int VW>
__global__ void __launch_bounds__(TB) inorm_bwd_reduce_kernel(const T* __restrict__ dy, const T* __restrict__ x,
                                                             const float* __restrict__ stats,
                                                             const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, int64_t voxels,
                                                             int C, float slope, int64_t rows_per_slab,
                                                             float* __restrict__ red) {
  extern __shared__ float smem[];
  const int n = blockIdx.y;
  const T* xb = x + (int64_t)n * voxels * C;
  const T* db = dy + (int64_t)n * voxels * C;
  const float* st = stats + (int64_t)n * C * 2;
  reduce_rows<2, VW>(voxels, C, rows_per_slab, smem, red + (int64_t)n * C * 2, 1.f,
                     [&](int64_t r, int cbase, float (&acc)[2][8]) {
                       float v[8], d[8], mean[8], rstd[8], g[8], b[8];
                       ldv<T, VW>(xb + r * C + cbase, v);
                       ldv<T, VW>(db + r * C + cbase, d);
                       ld_stats<VW>(st + cbase * 2, mean, rstd);
                       ldp<VW>(gamma + cbase, g);
                       ldp<VW>(beta + cbase, b);
#pragma unroll
                       for (int i = 0; i < VW; ++i) {
                         const float xh = (v[i] - mean[i]) * rstd[i];
                         const float gg = d[i] * ((xh * g[i] + b[i]) > 0.f ? 1.f : slope);
                         acc[0][i] += gg;
                         acc[1][i] = fmaf(gg, xh, acc[1][i]);
                       }
                     });
}

template <typename T, int VW>
__global__ void __launch_bounds__(TB) inorm_bwd_apply_kernel(const T* __restrict__ dy, const T* __restrict__ x,
                                                            const float* __restrict__ stats,
                                                            const float* __restrict__ gamma,
                                                            const float* __restrict__ beta,
                                                            const float* __restrict__ red, int64_t voxels, int C,
                                                            float slope, float inv_v, T* __restrict__ dx,
                                                            int accumulate, int64_t total_vec) {
  const int CG = C / VW;
  for (int64_t i = blockIdx.x * (int64_t)TB + threadIdx.x; i < total_vec; i += (int64_t)gridDim.x * TB) {
    const int cbase = (int)(i % CG) * VW;
    const int n = (int)(i / ((int64_t)CG * voxels));
    float v[8], d[8], mean[8], rstd[8], g[8], b[8], o[8];
    ldv<T, VW>(x + i * VW, v);
    ldv<T, VW>(dy + i * VW, d);
    ld_stats<VW>(stats + ((int64_t)n * C + cbase) * 2, mean, rstd);
    ldp<VW>(gamma + cbase, g);
    ldp<VW>(beta + cbase, b);
    const float* rd = red + ((int64_t)n * C + cbase) * 2;
    if (accumulate) ldv<T, VW>(dx + i * VW, o);
#pragma unroll
    for (int k = 0; k < VW; ++k) {
      const float xh = (v[k] - mean[k]) * rstd[k];
      const float gg = d[k] * ((xh * g[k] + b[k]) > 0.f ? 1.f : slope);
      const float r = g[k] * rstd[k] * (gg - rd[2 * k] * inv_v - xh * rd[2 * k + 1] * inv_v);
      o[k] = accumulate ? o[k] + r : r;
    }
    stv<T, VW>(dx + i * VW, o);
  }
}

// dgamma[c] += sum_n red[n][c][ig] (+extra), dbeta[c] += sum_n red[n][c][ib] (+ sum_n extra_b[n][c])
__global__ void param_grad_kernel(const float* __restrict__ red, int K, int ig, int ib, int batch, int C,
                                  const float* __restrict__ extra_b, float* __restrict__ dgamma,
                                  float* __restrict__ dbeta) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float sg = 0.f, sb = 0.f;
  for (int n = 0; n < batch; ++n) {
    sg += red[((int64_t)n * C + c) * K + ig];
    sb += red[((int64_t)n * C + c) * K + ib];
    if (extra_b) sb += extra_b[(int64_t)n * C + c];
  }
  if (dgamma) dgamma[c] += sg;
  if (dbeta) dbeta[c] += sb;
}

// ---------------------------------------------------------------------------------------------
// K5: squeeze / excite / gate
// ---------------------------------------------------------------------------------------------
template <typename T, int VW>
__global__ void __launch_bounds__(TB) se_squeeze_kernel(const T* __restrict__ raw3, const float* __restrict__ stats3,
                                                       const float* __restrict__ gamma3,
                                                       const float* __restrict__ beta3, int64_t voxels, int C,
                                                       int64_t rows_per_slab, float inv_v,
                                                       float* __restrict__ pool) {
  extern __shared__ float smem[];
  const int n = blockIdx.y;
  const T* xb = raw3 + (int64_t)n * voxels * C;
  const float* st = stats3 + (int64_t)n * C * 2;
  reduce_rows<1, VW>(voxels, C, rows_per_slab, smem, pool + (int64_t)n * C, inv_v,
                     [&](int64_t r, int cbase, float (&acc)[1][8]) {
                       float v[8], mean[8], rstd[8], g[8], b[8];
                       ldv<T, VW>(xb + r * C + cbase, v);
                       ld_stats<VW>(st + cbase * 2, mean, rstd);
                       ldp<VW>(gamma3 + cbase, g);
                       ldp<VW>(beta3 + cbase, b);
#pragma unroll
                       for (int i = 0; i < VW; ++i) acc[0][i] += (v[i] - mean[i]) * rstd[i] * g[i] + b[i];
                     });
}

// one block per sample; C <= 2048, Cr <= 256
__global__ void __launch_bounds__(TB) se_excite_fwd_kernel(const float* __restrict__ pool, const float* __restrict__ w6,
                                                          const float* __restrict__ b6, const float* __restrict__ w7,
                                                          const float* __restrict__ b7, int C, int Cr,
                                                          float* __restrict__ hidden, float* __restrict__ gate) {
  extern __shared__ float sm[];  // pool[C] | act[Cr]
  float* sp = sm;
  float* sa = sm + C;
  const int n = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += TB) sp[c] = pool[(int64_t)n * C + c];
  __syncthreads();
  for (int j = threadIdx.x; j < Cr; j += TB) {
    float h = b6[j];
    for (int c = 0; c < C; ++c) h = fmaf(sp[c], w6[(int64_t)c * Cr + j], h);
    hidden[(int64_t)n * Cr + j] = h;
    sa[j] = lrelu(h, M1_LRELU_SLOPE);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += TB) {
    float s = b7[c];
    for (int j = 0; j < Cr; ++j) s = fmaf(sa[j], w7[(int64_t)j * C + c], s);
    gate[(int64_t)n * C + c] = 1.f / (1.f + __expf(-s));
  }
}

__global__ void __launch_bounds__(TB) se_excite_bwd_kernel(const float* __restrict__ dgate, const float* __restrict__ pool,
                                                          const float* __restrict__ hidden,
                                                          const float* __restrict__ gate,
                                                          const float* __restrict__ w6, const float* __restrict__ w7,
                                                          int C, int Cr, float* __restrict__ dpool,
                                                          float* __restrict__ dw6, float* __restrict__ db6,
                                                          float* __restrict__ dw7, float* __restrict__ db7) {
  extern __shared__ float sm[];  // dpre7[C] | act[Cr] | dhid[Cr] | pool[C]
  float* sd7 = sm;
  float* sa = sm + C;
  float* sdh = sa + Cr;
  float* sp = sdh + Cr;
  const int n = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += TB) {
    const float g = gate[(int64_t)n * C + c];
    const float d = dgate[(int64_t)n * C + c] * g * (1.f - g);
    sd7[c] = d;
    sp[c] = pool[(int64_t)n * C + c];
    atomicAdd(&db7[c], d);
  }
  for (int j = threadIdx.x; j < Cr; j += TB) sa[j] = lrelu(hidden[(int64_t)n * Cr + j], M1_LRELU_SLOPE);
  __syncthreads();
  for (int j = threadIdx.x; j < Cr; j += TB) {
    float da = 0.f;
    for (int c = 0; c < C; ++c) da = fmaf(w7[(int64_t)j * C + c], sd7[c], da);
    const float dh = da * (hidden[(int64_t)n * Cr + j] > 0.f ? 1.f : M1_LRELU_SLOPE);
    sdh[j] = dh;
    atomicAdd(&db6[j], dh);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C * Cr; i += TB) {
    const int j7 = i / C, c7 = i % C;           // dW7[j][c] += act[j] * dpre7[c]
    atomicAdd(&dw7[i], sa[j7] * sd7[c7]);
    const int c6 = i / Cr, j6 = i % Cr;         // dW6[c][j] += pool[c] * dhid[j]
    atomicAdd(&dw6[i], sp[c6] * sdh[j6]);
  }
  for (int c = threadIdx.x; c < C; c += TB) {
    float dp = 0.f;
    for (int j = 0; j < Cr; ++j) dp = fmaf(w6[(int64_t)c * Cr + j], sdh[j], dp);
    dpool[(int64_t)n * C + c] = dp;
  }
}

struct DropArgs {
  const float* u;
  const uint64_t* step;          // device step counter of a replayed CUDA graph (or NULL)
  uint64_t seed, stream_id;
  float rate, scale;
};
// Philox stream of this launch: m1_dropout.stream_id, advanced by the device-resident step counter when the
// launch is part of a captured graph (scalar kernel arguments are frozen at capture time)
__device__ __forceinline__ uint64_t drop_stream(const DropArgs& dr) {
  return dr.stream_id + (dr.step ? *dr.step * M1_PHILOX_STEP_STRIDE : 0ull);
}

// keep-mask * scale for the VW elements starting at flat element index e (e % VW == 0)
template <int VW>
__device__ __forceinline__ void drop_factors(const DropArgs& dr, int64_t e, float (&f)[8]) {
  if (dr.rate <= 0.f) {
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = 1.f;
    return;
  }
  float u[8];
  if (dr.u != nullptr) {
    ldp<VW>(dr.u + e, u);
  } else {
    float q[4];
    const uint64_t sid = drop_stream(dr);
    philox_uniform4(dr.seed, sid, (uint64_t)(e >> 2), q);
    if constexpr (VW >= 4) {
      u[0] = q[0]; u[1] = q[1]; u[2] = q[2]; u[3] = q[3];
      if constexpr (VW == 8) {
        philox_uniform4(dr.seed, sid, (uint64_t)(e >> 2) + 1, q);
        u[4] = q[0]; u[5] = q[1]; u[6] = q[2]; u[7] = q[3];
      }
    } else {
      u[0] = q[e & 3];
    }
  }
#pragma unroll
  for (int i = 0; i < VW; ++i) f[i] = u[i] >= dr.rate ? dr.scale : 0.f;
}

struct GateArgs {
  const float *stats3, *stats4, *gamma3, *beta3, *gamma4, *beta4, *gate;
};

template <typename T, int VW>
__global__ void __launch_bounds__(TB) se_gate_fwd_kernel(const T* __restrict__ raw3, const T* __restrict__ raw4,
                                                        GateArgs a, DropArgs dr, int64_t voxels, int C,
                                                        T* __restrict__ out, int64_t total_vec) {
  const int CG = C / VW;
  for (int64_t i = blockIdx.x * (int64_t)TB + threadIdx.x; i < total_vec; i += (int64_t)gridDim.x * TB) {
    const int cbase = (int)(i % CG) * VW;
    const int n = (int)(i / ((int64_t)CG * voxels));
    const int64_t nc = (int64_t)n * C + cbase;
    float x3[8], x4[8], m3[8], r3[8], m4[8], r4[8], g3[8], b3[8], g4[8], b4[8], gt[8], f[8], o[8];
    ldv<T, VW>(raw3 + i * VW, x3);
    ldv<T, VW>(raw4 + i * VW, x4);
    ld_stats<VW>(a.stats3 + nc * 2, m3, r3);
    ld_stats<VW>(a.stats4 + nc * 2, m4, r4);
    ldp<VW>(a.gamma3 + cbase, g3); ldp<VW>(a.beta3 + cbase, b3);
    ldp<VW>(a.gamma4 + cbase, g4); ldp<VW>(a.beta4 + cbase, b4);
    ldp<VW>(a.gate + nc, gt);
    drop_factors<VW>(dr, i * VW, f);
#pragma unroll
    for (int k = 0; k < VW; ++k) {
      const float x_ = (x3[k] - m3[k]) * r3[k] * g3[k] + b3[k];
      const float res = (x4[k] - m4[k]) * r4[k] * g4[k] + b4[k];
      o[k] = lrelu(x_ * gt[k] * res, M1_LRELU_SLOPE) * f[k];
    }
    stv<T, VW>(out + i * VW, o);
  }
}

// red[n][c] = { sum dx_, sum dx_*xh3, sum dr, sum dr*xh4 } ; dgate[n][c] = sum dz*x_*r
template <typename T, int VW>
__global__ void __launch_bounds__(TB) se_gate_bwd_reduce_kernel(const T* __restrict__ dout, const T* __restrict__ raw3,
                                                               const T* __restrict__ raw4, GateArgs a, DropArgs dr,
                                                               int64_t voxels, int C, int64_t rows_per_slab,
                                                               float* __restrict__ red5) {
  extern __shared__ float smem[];
  const int n = blockIdx.y;
  const int64_t base = (int64_t)n * voxels * C;
  reduce_rows<5, VW>(voxels, C, rows_per_slab, smem, red5 + (int64_t)n * C * 5, 1.f,
                     [&](int64_t r, int cbase, float (&acc)[5][8]) {
                       const int64_t e = base + r * C + cbase;
                       const int64_t nc = (int64_t)n * C + cbase;
                       float x3[8], x4[8], d[8], m3[8], r3[8], m4[8], r4[8], g3[8], b3[8], g4[8], b4[8], gt[8], f[8];
                       ldv<T, VW>(raw3 + e, x3);
                       ldv<T, VW>(raw4 + e, x4);
                       ldv<T, VW>(dout + e, d);
                       ld_stats<VW>(a.stats3 + nc * 2, m3, r3);
                       ld_stats<VW>(a.stats4 + nc * 2, m4, r4);
                       ldp<VW>(a.gamma3 + cbase, g3); ldp<VW>(a.beta3 + cbase, b3);
                       ldp<VW>(a.gamma4 + cbase, g4); ldp<VW>(a.beta4 + cbase, b4);
                       ldp<VW>(a.gate + nc, gt);
                       drop_factors<VW>(dr, e, f);
#pragma unroll
                       for (int k = 0; k < VW; ++k) {
                         const float xh3 = (x3[k] - m3[k]) * r3[k], xh4 = (x4[k] - m4[k]) * r4[k];
                         const float x_ = xh3 * g3[k] + b3[k], res = xh4 * g4[k] + b4[k];
                         const float z = x_ * gt[k] * res;
                         const float dz = d[k] * f[k] * (z > 0.f ? 1.f : M1_LRELU_SLOPE);
                         const float dx_ = dz * gt[k] * res, dres = dz * x_ * gt[k];
                         acc[0][k] += dx_;
                         acc[1][k] = fmaf(dx_, xh3, acc[1][k]);
                         acc[2][k] += dres;
                         acc[3][k] = fmaf(dres, xh4, acc[3][k]);
                         acc[4][k] = fmaf(dz * x_, res, acc[4][k]);
                       }
                     });
}

__global__ void extract_dgate_kernel(const float* __restrict__ red5, int total, float* __restrict__ dgate) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < total) dgate[i] = red5[(int64_t)i * 5 + 4];
}

template <typename T, int VW>
__global__ void __launch_bounds__(TB) se_gate_bwd_apply_kernel(const T* __restrict__ dout, const T* __restrict__ raw3,
                                                              const T* __restrict__ raw4, GateArgs a, DropArgs dr,
                                                              const float* __restrict__ red5, int64_t voxels, int C,
                                                              float inv_v, T* __restrict__ draw3,
                                                              T* __restrict__ draw4, int64_t total_vec) {
  const int CG = C / VW;
  for (int64_t i = blockIdx.x * (int64_t)TB + threadIdx.x; i < total_vec; i += (int64_t)gridDim.x * TB) {
    const int cbase = (int)(i % CG) * VW;
    const int n = (int)(i / ((int64_t)CG * voxels));
    const int64_t nc = (int64_t)n * C + cbase;
    float x3[8], x4[8], d[8], m3[8], r3[8], m4[8], r4[8], g3[8], b3[8], g4[8], b4[8], gt[8], f[8], o3[8], o4[8];
    ldv<T, VW>(raw3 + i * VW, x3);
    ldv<T, VW>(raw4 + i * VW, x4);
    ldv<T, VW>(dout + i * VW, d);
    ld_stats<VW>(a.stats3 + nc * 2, m3, r3);
    ld_stats<VW>(a.stats4 + nc * 2, m4, r4);
    ldp<VW>(a.gamma3 + cbase, g3); ldp<VW>(a.beta3 + cbase, b3);
    ldp<VW>(a.gamma4 + cbase, g4); ldp<VW>(a.beta4 + cbase, b4);
    ldp<VW>(a.gate + nc, gt);
    drop_factors<VW>(dr, i * VW, f);
    const float* rd = red5 + nc * 5;
#pragma unroll
    for (int k = 0; k < VW; ++k) {
      const float xh3 = (x3[k] - m3[k]) * r3[k], xh4 = (x4[k] - m4[k]) * r4[k];
      const float x_ = xh3 * g3[k] + b3[k], res = xh4 * g4[k] + b4[k];
      const float z = x_ * gt[k] * res;
      const float dz = d[k] * f[k] * (z > 0.f ? 1.f : M1_LRELU_SLOPE);
      const float dx_ = dz * gt[k] * res, dres = dz * x_ * gt[k];
      // the GAP path adds dpool/V to dx_, a per-(n,c) constant that the norm backward removes again
      o3[k] = g3[k] * r3[k] * (dx_ - rd[5 * k + 0] * inv_v - xh3 * rd[5 * k + 1] * inv_v);
      o4[k] = g4[k] * r4[k] * (dres - rd[5 * k + 2] * inv_v - xh4 * rd[5 * k + 3] * inv_v);
    }
    stv<T, VW>(draw3 + i * VW, o3);
    stv<T, VW>(draw4 + i * VW, o4);
  }
}

inline int64_t slab_rows(const m1_ctx* ctx, int batch, int64_t voxels) {
  // ~4 blocks per SM over the whole launch
  int64_t slabs = std::max<int64_t>(1, (int64_t)ctx->num_sms * 4 / std::max(1, batch));
  int64_t rows = cdiv64(voxels, slabs);
  return std::max<int64_t>(rows, 32);
}
inline unsigned ew_blocks(const m1_ctx* ctx, int64_t total_vec) {
  return (unsigned)std::max<int64_t>(1, std::min<int64_t>(cdiv64(total_vec, TB), (int64_t)ctx->num_sms * 16));
}

#define DISPATCH_VW_(C, ...)                                             \
  do {                                                                   \
    if ((C) % 8 == 0) { constexpr int VW = 8; __VA_ARGS__; }             \
    else if ((C) % 4 == 0) { constexpr int VW = 4; __VA_ARGS__; }        \
    else { constexpr int VW = 1; __VA_ARGS__; }                          \
  } while (0)
#define DISPATCH_VW4_(C, ...)                                            \
  do {                                                                   \
    if ((C) % 4 == 0) { constexpr int VW = 4; __VA_ARGS__; }             \
    else { constexpr int VW = 1; __VA_ARGS__; }                          \
  } while (0)
#define DISPATCH_T_VW4(dtype, C, ...)                                    \
  do {                                                                   \
    if ((dtype) == M1_BF16) {                                            \
      using T = __nv_bfloat16;                                           \
      DISPATCH_VW4_(C, __VA_ARGS__);                                     \
    } else {                                                             \
      using T = float;                                                   \
      DISPATCH_VW4_(C, __VA_ARGS__);                                     \
    }                                                                    \
  } while (0)
#define DISPATCH_T_VW(dtype, C, ...)                                     \
  do {                                                                   \
    if ((dtype) == M1_BF16) {                                            \
      using T = __nv_bfloat16;                                           \
      DISPATCH_VW_(C, __VA_ARGS__);                                      \
    } else {                                                             \
      using T = float;                                                   \
      DISPATCH_VW_(C, __VA_ARGS__);                                      \
    }                                                                    \
  } while (0)

DropArgs make_drop(const m1_dropout* d) {
  DropArgs a;
  a.u = d ? d->u : nullptr;
  a.seed = d ? d->seed : 0;
  a.stream_id = d ? d->stream_id : 0;
  a.step = d ? d->step : nullptr;
  a.rate = d ? d->rate : 0.f;
  a.scale = (d && d->rate > 0.f) ? 1.f / (1.f - d->rate) : 1.f;
  return a;
}

}  // namespace

extern "C" int m1_inorm_stats(m1_ctx* ctx, const void* x, int dtype, int batch, int64_t voxels, int C,
                              float eps, float* stats, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  M1_CUDA(cudaMemsetAsync(stats, 0, (size_t)batch * C * 2 * sizeof(float), st));
  const int64_t rows = slab_rows(ctx, batch, voxels);
  dim3 grid((unsigned)cdiv64(voxels, rows), (unsigned)batch);
  DISPATCH_T_VW4(dtype, C, (inorm_sums_kernel<T, VW><<<grid, TB, 2 * C * sizeof(float), st>>>(
                              reinterpret_cast<const T*>(x), voxels, C, rows, stats)));
  M1_LAUNCH_CHECK(ctx);
  const int total = batch * C;
  inorm_finalize_kernel<<<(total + 255) / 256, 256, 0, st>>>(stats, total, 1.f / (float)voxels, eps);
  M1_LAUNCH_CHECK(ctx);
  return 0;
}

extern "C" int m1_inorm_act_fwd(m1_ctx* ctx, const void* x, const float* stats, const float* gamma,
                                const float* beta, int dtype, int batch, int64_t voxels, int C, float slope,
                                void* y, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  DISPATCH_T_VW(dtype, C, {
    const int64_t total_vec = (int64_t)batch * voxels * (C / VW);
    inorm_act_fwd_kernel<T, VW><<<ew_blocks(ctx, total_vec), TB, 0, st>>>(
        reinterpret_cast<const T*>(x), stats, gamma, beta, voxels, C, slope, reinterpret_cast<T*>(y), total_vec);
  });
  M1_LAUNCH_CHECK(ctx);
  return 0;
}

extern "C" int m1_inorm_act_bwd(m1_ctx* ctx, const void* dy, const void* x, const float* stats,
                                const float* gamma, const float* beta, int dtype, int batch, int64_t voxels,
                                int C, float slope, void* dx, int accumulate, float* dgamma, float* dbeta,
                                void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  M1_CHECK((size_t)batch * C * 2 * sizeof(float) <= ctx->scratch_bytes, "inorm_act_bwd: scratch too small");
  float* red = ctx->scratch;
  M1_CUDA(cudaMemsetAsync(red, 0, (size_t)batch * C * 2 * sizeof(float), st));
  const int64_t rows = slab_rows(ctx, batch, voxels);
  dim3 grid((unsigned)cdiv64(voxels, rows), (unsigned)batch);
  DISPATCH_T_VW4(dtype, C, (inorm_bwd_reduce_kernel<T, VW><<<grid, TB, 2 * C * sizeof(float), st>>>(
                              reinterpret_cast<const T*>(dy), reinterpret_cast<const T*>(x), stats, gamma, beta,
                              voxels, C, slope, rows, red)));
  M1_LAUNCH_CHECK(ctx);
  DISPATCH_T_VW(dtype, C, {
    const int64_t total_vec = (int64_t)batch * voxels * (C / VW);
    inorm_bwd_apply_kernel<T, VW><<<ew_blocks(ctx, total_vec), TB, 0, st>>>(
        reinterpret_cast<const T*>(dy), reinterpret_cast<const T*>(x), stats, gamma, beta, red, voxels, C, slope,
        1.f / (float)voxels, reinterpret_cast<T*>(dx), accumulate, total_vec);
  });
  M1_LAUNCH_CHECK(ctx);
  if (dgamma || dbeta) {
    param_grad_kernel<<<(C + 127) / 128, 128, 0, st>>>(red, 2, 1, 0, batch, C, nullptr, dgamma, dbeta);
    M1_LAUNCH_CHECK(ctx);
  }
  return 0;
}

extern "C" int m1_se_squeeze(m1_ctx* ctx, const void* raw3, const float* stats3, const float* gamma3,
                             const float* beta3, int dtype, int batch, int64_t voxels, int C, float* pool,
                             void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  M1_CUDA(cudaMemsetAsync(pool, 0, (size_t)batch * C * sizeof(float), st));
  const int64_t rows = slab_rows(ctx, batch, voxels);
  dim3 grid((unsigned)cdiv64(voxels, rows), (unsigned)batch);
  DISPATCH_T_VW4(dtype, C, (se_squeeze_kernel<T, VW><<<grid, TB, C * sizeof(float), st>>>(
                              reinterpret_cast<const T*>(raw3), stats3, gamma3, beta3, voxels, C, rows,
                              1.f / (float)voxels, pool)));
  M1_LAUNCH_CHECK(ctx);
  return 0;
}

extern "C" int m1_se_excite_fwd(m1_ctx* ctx, const float* pool, const float* w6, const float* b6,
                                const float* w7, const float* b7, int batch, int C, int Cr, float* hidden,
                                float* gate, void* stream) {
  se_excite_fwd_kernel<<<batch, TB, (C + Cr) * sizeof(float), (cudaStream_t)stream>>>(pool, w6, b6, w7, b7, C, Cr,
                                                                                     hidden, gate);
  M1_LAUNCH_CHECK(ctx);
  return 0;
}

extern "C" int m1_se_excite_bwd(m1_ctx* ctx, const float* dgate, const float* pool, const float* hidden,
                                const float* gate, const float* w6, const float* w7, int batch, int C, int Cr,
                                float* dpool, float* dw6, float* db6, float* dw7, float* db7, void* stream) {
  se_excite_bwd_kernel<<<batch, TB, (2 * C + 2 * Cr) * sizeof(float), (cudaStream_t)stream>>>(
      dgate, pool, hidden, gate, w6, w7, C, Cr, dpool, dw6, db6, dw7, db7);
  M1_LAUNCH_CHECK(ctx);
  return 0;
}

extern "C" int m1_se_gate_fwd(m1_ctx* ctx, const void* raw3, const void* raw4, const float* stats3,
                              const float* stats4, const float* gamma3, const float* beta3, const float* gamma4,
                              const float* beta4, const float* gate, const m1_dropout* drop, int dtype, int batch,
                              int64_t voxels, int C, void* out, void* stream) {
  GateArgs a{stats3, stats4, gamma3, beta3, gamma4, beta4, gate};
  DropArgs dr = make_drop(drop);
  DISPATCH_T_VW(dtype, C, {
    const int64_t total_vec = (int64_t)batch * voxels * (C / VW);
    se_gate_fwd_kernel<T, VW><<<ew_blocks(ctx, total_vec), TB, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const T*>(raw3), reinterpret_cast<const T*>(raw4), a, dr, voxels, C,
        reinterpret_cast<T*>(out), total_vec);
  });
  M1_LAUNCH_CHECK(ctx);
  return 0;
}

extern "C" int m1_se_gate_bwd_reduce(m1_ctx* ctx, const void* dout, const void* raw3, const void* raw4,
                                     const float* stats3, const float* stats4, const float* gamma3,
                                     const float* beta3, const float* gamma4, const float* beta4,
                                     const float* gate, const m1_dropout* drop, int dtype, int batch,
                                     int64_t voxels, int C, float* red, float* dgate, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  GateArgs a{stats3, stats4, gamma3, beta3, gamma4, beta4, gate};
  DropArgs dr = make_drop(drop);
  M1_CUDA(cudaMemsetAsync(red, 0, (size_t)batch * C * 5 * sizeof(float), st));
  const int64_t rows = slab_rows(ctx, batch, voxels);
  dim3 grid((unsigned)cdiv64(voxels, rows), (unsigned)batch);
  DISPATCH_T_VW4(dtype, C, (se_gate_bwd_reduce_kernel<T, VW><<<grid, TB, 5 * C * sizeof(float), st>>>(
                              reinterpret_cast<const T*>(dout), reinterpret_cast<const T*>(raw3),
                              reinterpret_cast<const T*>(raw4), a, dr, voxels, C, rows, red)));
  M1_LAUNCH_CHECK(ctx);
  const int total = batch * C;
  extract_dgate_kernel<<<(total + 255) / 256, 256, 0, st>>>(red, total, dgate);
  M1_LAUNCH_CHECK(ctx);
  return 0;
}

extern "C" int m1_se_gate_bwd_apply(m1_ctx* ctx, const void* dout, const void* raw3, const void* raw4,
                                    const float* stats3, const float* stats4, const float* gamma3,
                                    const float* beta3, const float* gamma4, const float* beta4,
                                    const float* gate, const m1_dropout* drop, const float* red,
                                    const float* dpool, int dtype, int batch, int64_t voxels, int C, void* draw3,
                                    void* draw4, float* dgamma3, float* dbeta3, float* dgamma4, float* dbeta4,
                                    void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  GateArgs a{stats3, stats4, gamma3, beta3, gamma4, beta4, gate};
  DropArgs dr = make_drop(drop);
  DISPATCH_T_VW(dtype, C, {
    const int64_t total_vec = (int64_t)batch * voxels * (C / VW);
    se_gate_bwd_apply_kernel<T, VW><<<ew_blocks(ctx, total_vec), TB, 0, st>>>(
        reinterpret_cast<const T*>(dout), reinterpret_cast<const T*>(raw3), reinterpret_cast<const T*>(raw4), a, dr,
        red, voxels, C, 1.f / (float)voxels, reinterpret_cast<T*>(draw3), reinterpret_cast<T*>(draw4), total_vec);
  });
  M1_LAUNCH_CHECK(ctx);
  // norm3: dgamma += sum_n A2, dbeta += sum_n (A1 + dpool) ; norm4: dgamma += sum_n B2, dbeta += sum_n B1
  param_grad_kernel<<<(C + 127) / 128, 128, 0, st>>>(red, 5, 1, 0, batch, C, dpool, dgamma3, dbeta3);
  M1_LAUNCH_CHECK(ctx);
  param_grad_kernel<<<(C + 127) / 128, 128, 0, st>>>(red, 5, 3, 2, batch, C, nullptr, dgamma4, dbeta4);
  M1_LAUNCH_CHECK(ctx);
  return 0;
}
