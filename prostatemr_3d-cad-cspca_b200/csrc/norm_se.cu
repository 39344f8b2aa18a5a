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

// Raw (unconverted) vector of VW channels: what a load returns before anything depends on it. The skeletons
// below first issue the loads of UNR rows x NT tensors into RawVec registers and only then start converting
// and computing - explicit memory-level parallelism (these kernels are latency bound on long-scoreboard
// stalls otherwise: ~50 KB must be in flight per SM to saturate HBM3e).
template <typename T, int VW>
struct RawVec {
  static constexpr int kWords = (int)(sizeof(T) * VW + 3) / 4;
  uint32_t w[kWords];
  __device__ __forceinline__ void load(const T* p) {
    if constexpr (sizeof(T) * VW == 32) {
      const uint4 a = *reinterpret_cast<const uint4*>(p), b = *(reinterpret_cast<const uint4*>(p) + 1);
      w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
    } else if constexpr (sizeof(T) * VW == 16) {
      const uint4 a = *reinterpret_cast<const uint4*>(p);
      w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w;
    } else if constexpr (sizeof(T) * VW == 8) {
      const uint2 a = *reinterpret_cast<const uint2*>(p);
      w[0] = a.x; w[1] = a.y;
    } else if constexpr (sizeof(T) == 4) {
      w[0] = *reinterpret_cast<const uint32_t*>(p);
    } else {
      w[0] = *reinterpret_cast<const uint16_t*>(p);
    }
  }
  __device__ __forceinline__ void unpack(float (&v)[8]) const {
    if constexpr (sizeof(T) == 4) {
#pragma unroll
      for (int i = 0; i < VW; ++i) v[i] = __uint_as_float(w[i]);
    } else if constexpr (VW == 1) {
      v[0] = __uint_as_float(w[0] << 16);
    } else {
#pragma unroll
      for (int i = 0; i < VW / 2; ++i) {
        v[2 * i] = __uint_as_float(w[i] << 16);            // bf16 -> fp32: the bits are the high half
        v[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
      }
    }
  }
};

// Row-parallel skeletons. grid = (slabs, batch); a thread owns ONE channel group (VW channels) and a row
// lane: the per-channel parameters (statistics, affine, gate, reduction results) are loaded and folded ONCE
// into registers by `prep(cbase)`, then the thread streams rows [slab*rows_per_slab, ...) with stride
// `lanes` - one vector access per tensor and row, no index arithmetic, no parameter traffic in the loop.
// Consecutive threads cover consecutive channel groups of a row, then the next row: fully coalesced.
// src[t]: the NT streamed tensors of this sample; F gets the raw vectors of one row.
//
// reduce_rows: F(row, cbase, regs, raw[NT], acc[K][8]) accumulates; the block folds its partial sums through
// warp shuffles + shared memory and issues one global atomic per (block, channel, k).
template <int K, int VW, int UNR, int NT, typename T, typename P, typename F>
__device__ __forceinline__ void reduce_rows(const T* const (&src)[NT], int64_t voxels, int C, int64_t rows_per_slab,
                                            float* smem, float* gout /* [C][K] of this sample */, float scale,
                                            P prep, F f) {
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
      const auto regs = prep(cbase);
      const int64_t step = (int64_t)lanes * C;
      int64_t r = r0 + my_lane;
      int64_t off = r * C + cbase;
      for (; r + (int64_t)(UNR - 1) * lanes < r1; r += (int64_t)UNR * lanes, off += UNR * step) {
        RawVec<T, VW> raw[UNR][NT];
#pragma unroll
        for (int u = 0; u < UNR; ++u)
#pragma unroll
          for (int t = 0; t < NT; ++t) raw[u][t].load(src[t] + off + u * step);
#pragma unroll
        for (int u = 0; u < UNR; ++u) f(r + (int64_t)u * lanes, cbase, regs, raw[u], acc);
      }
      for (; r < r1; r += lanes, off += step) {
        RawVec<T, VW> raw[NT];
#pragma unroll
        for (int t = 0; t < NT; ++t) raw[t].load(src[t] + off);
        f(r, cbase, regs, raw, acc);
      }
      // lanes of a warp that own the same channel group (cgs < 32: lane ids congruent mod cgs) fold their
      // partial sums with shuffles first: one shared-memory atomic per (warp, channel, k) instead of per thread
      const bool fold = cgs < 32 && (cgs & (cgs - 1)) == 0;     // power of two: every thread of the block is active
      if (fold) {
#pragma unroll
        for (int k = 0; k < K; ++k)
#pragma unroll
          for (int i = 0; i < VW; ++i)
            for (int o = cgs; o < 32; o <<= 1) acc[k][i] += __shfl_xor_sync(0xffffffffu, acc[k][i], o);
      }
      if (!fold || (threadIdx.x & 31) < cgs) {
#pragma unroll
        for (int k = 0; k < K; ++k)
#pragma unroll
          for (int i = 0; i < VW; ++i) atomicAdd(&smem[k * C + cbase + i], acc[k][i]);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < K * C; i += TB) {
    const int k = i / C, c = i % C;
    atomicAdd(&gout[c * K + k], smem[i] * scale);
  }
}

// stream_rows: pure elementwise pass, F(row, element offset, cbase, regs, raw[NT]) computes and stores one row's
// channel group (element offset = row * C + cbase inside the sample)
template <int VW, int UNR, int NT, typename T, typename P, typename F>
__device__ __forceinline__ void stream_rows(const T* const (&src)[NT], int64_t voxels, int C, int64_t rows_per_slab,
                                            P prep, F f) {
  const int CG = C / VW;
  const int cgs = min(CG, TB);
  const int lanes = TB / cgs;
  const int my_lane = threadIdx.x / cgs;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_slab;
  const int64_t r1 = min(r0 + rows_per_slab, voxels);
  if (my_lane >= lanes) return;
  for (int cg = threadIdx.x % cgs; cg < CG; cg += cgs) {
    const int cbase = cg * VW;
    const auto regs = prep(cbase);
    const int64_t step = (int64_t)lanes * C;
    int64_t r = r0 + my_lane;
    int64_t off = r * C + cbase;
    for (; r + (int64_t)(UNR - 1) * lanes < r1; r += (int64_t)UNR * lanes, off += UNR * step) {
      RawVec<T, VW> raw[UNR][NT];
#pragma unroll
      for (int u = 0; u < UNR; ++u)
#pragma unroll
        for (int t = 0; t < NT; ++t) raw[u][t].load(src[t] + off + u * step);
#pragma unroll
      for (int u = 0; u < UNR; ++u) f(r + (int64_t)u * lanes, off + u * step, cbase, regs, raw[u]);
    }
    for (; r < r1; r += lanes, off += step) {
      RawVec<T, VW> raw[NT];
#pragma unroll
      for (int t = 0; t < NT; ++t) raw[t].load(src[t] + off);
      f(r, off, cbase, regs, raw);
    }
  }
}

// instance norm folded to y = x * a + b  (a = rstd * gamma, b = beta - mean * a), plus what backward needs
struct NormRegs {
  float a[8], b[8], mean[8], rstd[8], gamma[8];
};
template <int VW>
__device__ __forceinline__ NormRegs norm_regs(const float* st /* [C][2] of the sample */, const float* gamma,
                                              const float* beta, int cbase) {
  NormRegs r;
  float g[8], b[8];
  ld_stats<VW>(st + cbase * 2, r.mean, r.rstd);
  ldp<VW>(gamma + cbase, g);
  ldp<VW>(beta + cbase, b);
#pragma unroll
  for (int i = 0; i < VW; ++i) {
    r.gamma[i] = g[i];
    r.a[i] = r.rstd[i] * g[i];
    r.b[i] = b[i] - r.mean[i] * r.a[i];
  }
  return r;
}

// ---------------------------------------------------------------------------------------------
// K4 forward
// ---------------------------------------------------------------------------------------------
template <typename T, int VW>
__global__ void __launch_bounds__(TB) inorm_sums_kernel(const T* __restrict__ x, int64_t voxels, int C,
                                                       int64_t rows_per_slab, float* __restrict__ sums) {
  extern __shared__ float smem[];
  const int n = blockIdx.y;
  const T* const src[1] = {x + (int64_t)n * voxels * C};
  reduce_rows<2, VW, 8, 1>(src, voxels, C, rows_per_slab, smem, sums + (int64_t)n * C * 2, 1.f, [](int) { return 0; },
                     [&](int64_t, int, int, const RawVec<T, VW> (&raw)[1], float (&acc)[2][8]) {
                       float v[8];
                       raw[0].unpack(v);
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
                                                          float slope, T* __restrict__ y, int64_t rows_per_slab) {
  const int n = blockIdx.y;
  const T* const src[1] = {x + (int64_t)n * voxels * C};
  T* yb = y + (int64_t)n * voxels * C;
  stream_rows<VW, 8, 1>(src, voxels, C, rows_per_slab,
                  [&](int cbase) { return norm_regs<VW>(stats + (int64_t)n * C * 2, gamma, beta, cbase); },
                  [&](int64_t, int64_t off, int, const NormRegs& q, const RawVec<T, VW> (&raw)[1]) {
                    float v[8];
                    raw[0].unpack(v);
#pragma unroll
                    for (int k = 0; k < VW; ++k) v[k] = lrelu(fmaf(v[k], q.a[k], q.b[k]), slope);
                    stv<T, VW>(yb + off, v);
                  });
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
  const T* const src[2] = {x + (int64_t)n * voxels * C, dy + (int64_t)n * voxels * C};
  reduce_rows<2, VW, 4, 2>(src, voxels, C, rows_per_slab, smem, red + (int64_t)n * C * 2, 1.f,
                     [&](int cbase) { return norm_regs<VW>(stats + (int64_t)n * C * 2, gamma, beta, cbase); },
                     [&](int64_t, int, const NormRegs& q, const RawVec<T, VW> (&raw)[2], float (&acc)[2][8]) {
                       float v[8], d[8];
                       raw[0].unpack(v);
                       raw[1].unpack(d);
#pragma unroll
                       for (int i = 0; i < VW; ++i) {
                         const float xh = (v[i] - q.mean[i]) * q.rstd[i];
                         const float gg = d[i] * (fmaf(v[i], q.a[i], q.b[i]) > 0.f ? 1.f : slope);
                         acc[0][i] += gg;
                         acc[1][i] = fmaf(gg, xh, acc[1][i]);
                       }
                     });
}

struct NormBwdRegs {
  NormRegs q;
  float c1[8], c2[8];      // (sum g) / V, (sum g * xhat) / V
};

template <typename T, int VW>
__global__ void __launch_bounds__(TB) inorm_bwd_apply_kernel(const T* __restrict__ dy, const T* __restrict__ x,
                                                            const float* __restrict__ stats,
                                                            const float* __restrict__ gamma,
                                                            const float* __restrict__ beta,
                                                            const float* __restrict__ red, int64_t voxels, int C,
                                                            float slope, float inv_v, T* __restrict__ dx,
                                                            int accumulate, int64_t rows_per_slab) {
  const int n = blockIdx.y;
  const int64_t base = (int64_t)n * voxels * C;
  const T* const src[2] = {x + base, dy + base};
  T* dxb = dx + base;
  stream_rows<VW, 4, 2>(src, voxels, C, rows_per_slab,
                  [&](int cbase) {
                    NormBwdRegs w;
                    w.q = norm_regs<VW>(stats + (int64_t)n * C * 2, gamma, beta, cbase);
                    const float* rd = red + ((int64_t)n * C + cbase) * 2;
#pragma unroll
                    for (int k = 0; k < VW; ++k) { w.c1[k] = rd[2 * k] * inv_v; w.c2[k] = rd[2 * k + 1] * inv_v; }
                    return w;
                  },
                  [&](int64_t, int64_t off, int, const NormBwdRegs& w, const RawVec<T, VW> (&raw)[2]) {
                    float v[8], d[8], o[8];
                    raw[0].unpack(v);
                    raw[1].unpack(d);
                    if (accumulate) ldv<T, VW>(dxb + off, o);
#pragma unroll
                    for (int k = 0; k < VW; ++k) {
                      const float xh = (v[k] - w.q.mean[k]) * w.q.rstd[k];
                      const float gg = d[k] * (fmaf(v[k], w.q.a[k], w.q.b[k]) > 0.f ? 1.f : slope);
                      const float t = w.q.a[k] * (gg - w.c1[k] - xh * w.c2[k]);
                      o[k] = accumulate ? o[k] + t : t;
                    }
                    stv<T, VW>(dxb + off, o);
                  });
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
  const T* const src[1] = {raw3 + (int64_t)n * voxels * C};
  reduce_rows<1, VW, 8, 1>(src, voxels, C, rows_per_slab, smem, pool + (int64_t)n * C, inv_v,
                     [&](int cbase) { return norm_regs<VW>(stats3 + (int64_t)n * C * 2, gamma3, beta3, cbase); },
                     [&](int64_t, int, const NormRegs& q, const RawVec<T, VW> (&raw)[1], float (&acc)[1][8]) {
                       float v[8];
                       raw[0].unpack(v);
#pragma unroll
                       for (int i = 0; i < VW; ++i) acc[0][i] += fmaf(v[i], q.a[i], q.b[i]);
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
  uint8_t* mask;                 // keep-bits, one byte per 8 elements: written by the forward gate kernel, read by
                                 // its two backward kernels instead of re-running Philox (NULL: regenerate)
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
template <int VW, bool kBackward = false>
__device__ __forceinline__ void drop_factors(const DropArgs& dr, int64_t e, float (&f)[8]) {
  if (dr.rate <= 0.f) {
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = 1.f;
    return;
  }
  if constexpr (kBackward && VW >= 4) {
    if (dr.mask != nullptr) {
      const uint32_t bits = (uint32_t)dr.mask[e >> 3] >> (e & 7);      // e % VW == 0
#pragma unroll
      for (int i = 0; i < VW; ++i) f[i] = ((bits >> i) & 1u) ? dr.scale : 0.f;
      return;
    }
  }
  float u[8];
  if (dr.u != nullptr) {
    ldp<VW>(dr.u + e, u);
  } else {
    // one Philox4x32-10 block per 8 elements: element e takes the 16-bit half (e & 1) of word (e >> 1) & 3 of
    // block e >> 3 (the mapping does not depend on VW, so every kernel of a layer regenerates the same mask)
    uint32_t w[4];
    philox_words4(dr.seed, drop_stream(dr), (uint64_t)(e >> 3), w);
    if constexpr (VW == 8) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        u[2 * i] = (float)(w[i] & 0xFFFFu) * (1.0f / 65536.0f);
        u[2 * i + 1] = (float)(w[i] >> 16) * (1.0f / 65536.0f);
      }
    } else if constexpr (VW == 4) {
      const int w0 = (int)((e >> 1) & 3);
      u[0] = (float)(w[w0] & 0xFFFFu) * (1.0f / 65536.0f);
      u[1] = (float)(w[w0] >> 16) * (1.0f / 65536.0f);
      u[2] = (float)(w[w0 + 1] & 0xFFFFu) * (1.0f / 65536.0f);
      u[3] = (float)(w[w0 + 1] >> 16) * (1.0f / 65536.0f);
    } else {
      const uint32_t ww = w[(e >> 1) & 3];
      u[0] = (float)((e & 1) ? (ww >> 16) : (ww & 0xFFFFu)) * (1.0f / 65536.0f);
    }
  }
#pragma unroll
  for (int i = 0; i < VW; ++i) f[i] = u[i] >= dr.rate ? dr.scale : 0.f;
  if constexpr (!kBackward && VW == 8) {
    if (dr.mask != nullptr) {
      uint32_t bits = 0;
#pragma unroll
      for (int i = 0; i < 8; ++i) bits |= (f[i] != 0.f ? 1u : 0u) << i;
      dr.mask[e >> 3] = (uint8_t)bits;
    }
  }
}

struct GateArgs {
  const float *stats3, *stats4, *gamma3, *beta3, *gamma4, *beta4, *gate;
};

struct GateRegs {
  NormRegs q3, q4;
  float gt[8];
};
template <int VW>
__device__ __forceinline__ GateRegs gate_regs(const GateArgs& a, int n, int C, int cbase) {
  GateRegs w;
  w.q3 = norm_regs<VW>(a.stats3 + (int64_t)n * C * 2, a.gamma3, a.beta3, cbase);
  w.q4 = norm_regs<VW>(a.stats4 + (int64_t)n * C * 2, a.gamma4, a.beta4, cbase);
  ldp<VW>(a.gate + (int64_t)n * C + cbase, w.gt);
  return w;
}

template <typename T, int VW>
__global__ void __launch_bounds__(TB) se_gate_fwd_kernel(const T* __restrict__ raw3, const T* __restrict__ raw4,
                                                        GateArgs a, DropArgs dr, int64_t voxels, int C,
                                                        T* __restrict__ out, int64_t rows_per_slab) {
  const int n = blockIdx.y;
  const int64_t base = (int64_t)n * voxels * C;
  const T* const src[2] = {raw3 + base, raw4 + base};
  stream_rows<VW, 4, 2>(src, voxels, C, rows_per_slab, [&](int cbase) { return gate_regs<VW>(a, n, C, cbase); },
                  [&](int64_t, int64_t off, int, const GateRegs& w, const RawVec<T, VW> (&raw)[2]) {
                    const int64_t e = base + off;
                    float x3[8], x4[8], f[8], o[8];
                    raw[0].unpack(x3);
                    raw[1].unpack(x4);
                    drop_factors<VW>(dr, e, f);
#pragma unroll
                    for (int k = 0; k < VW; ++k) {
                      const float x_ = fmaf(x3[k], w.q3.a[k], w.q3.b[k]);
                      const float res = fmaf(x4[k], w.q4.a[k], w.q4.b[k]);
                      o[k] = lrelu(x_ * w.gt[k] * res, M1_LRELU_SLOPE) * f[k];
                    }
                    stv<T, VW>(out + e, o);
                  });
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
  const T* const src[3] = {raw3 + base, raw4 + base, dout + base};
  reduce_rows<5, VW, 4, 3>(src, voxels, C, rows_per_slab, smem, red5 + (int64_t)n * C * 5, 1.f,
                     [&](int cbase) { return gate_regs<VW>(a, n, C, cbase); },
                     [&](int64_t r, int cbase, const GateRegs& w, const RawVec<T, VW> (&raw)[3], float (&acc)[5][8]) {
                       const int64_t e = base + r * C + cbase;
                       float x3[8], x4[8], d[8], f[8];
                       raw[0].unpack(x3);
                       raw[1].unpack(x4);
                       raw[2].unpack(d);
                       drop_factors<VW, true>(dr, e, f);
#pragma unroll
                       for (int k = 0; k < VW; ++k) {
                         const float xh3 = (x3[k] - w.q3.mean[k]) * w.q3.rstd[k];
                         const float xh4 = (x4[k] - w.q4.mean[k]) * w.q4.rstd[k];
                         const float x_ = fmaf(x3[k], w.q3.a[k], w.q3.b[k]), res = fmaf(x4[k], w.q4.a[k], w.q4.b[k]);
                         const float z = x_ * w.gt[k] * res;
                         const float dz = d[k] * f[k] * (z > 0.f ? 1.f : M1_LRELU_SLOPE);
                         const float dx_ = dz * w.gt[k] * res, dres = dz * x_ * w.gt[k];
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

struct GateBwdRegs {
  GateRegs w;
  float c[4][8];      // the four reduction results / V
};

template <typename T, int VW>
__global__ void __launch_bounds__(TB) se_gate_bwd_apply_kernel(const T* __restrict__ dout, const T* __restrict__ raw3,
                                                              const T* __restrict__ raw4, GateArgs a, DropArgs dr,
                                                              const float* __restrict__ red5, int64_t voxels, int C,
                                                              float inv_v, T* __restrict__ draw3,
                                                              T* __restrict__ draw4, int64_t rows_per_slab) {
  const int n = blockIdx.y;
  const int64_t base = (int64_t)n * voxels * C;
  const T* const src[3] = {raw3 + base, raw4 + base, dout + base};
  stream_rows<VW, 4, 3>(src, voxels, C, rows_per_slab,
                  [&](int cbase) {
                    GateBwdRegs g;
                    g.w = gate_regs<VW>(a, n, C, cbase);
                    const float* rd = red5 + ((int64_t)n * C + cbase) * 5;
#pragma unroll
                    for (int k = 0; k < VW; ++k)
#pragma unroll
                      for (int j = 0; j < 4; ++j) g.c[j][k] = rd[5 * k + j] * inv_v;
                    return g;
                  },
                  [&](int64_t, int64_t off, int, const GateBwdRegs& g, const RawVec<T, VW> (&raw)[3]) {
                    const GateRegs& w = g.w;
                    const int64_t e = base + off;
                    float x3[8], x4[8], d[8], f[8], o3[8], o4[8];
                    raw[0].unpack(x3);
                    raw[1].unpack(x4);
                    raw[2].unpack(d);
                    drop_factors<VW, true>(dr, e, f);
#pragma unroll
                    for (int k = 0; k < VW; ++k) {
                      const float xh3 = (x3[k] - w.q3.mean[k]) * w.q3.rstd[k];
                      const float xh4 = (x4[k] - w.q4.mean[k]) * w.q4.rstd[k];
                      const float x_ = fmaf(x3[k], w.q3.a[k], w.q3.b[k]), res = fmaf(x4[k], w.q4.a[k], w.q4.b[k]);
                      const float z = x_ * w.gt[k] * res;
                      const float dz = d[k] * f[k] * (z > 0.f ? 1.f : M1_LRELU_SLOPE);
                      const float dx_ = dz * w.gt[k] * res, dres = dz * x_ * w.gt[k];
                      // the GAP path adds dpool/V to dx_, a per-(n,c) constant that the norm backward removes again
                      o3[k] = w.q3.a[k] * (dx_ - g.c[0][k] - xh3 * g.c[1][k]);
                      o4[k] = w.q4.a[k] * (dres - g.c[2][k] - xh4 * g.c[3][k]);
                    }
                    stv<T, VW>(draw3 + e, o3);
                    stv<T, VW>(draw4 + e, o4);
                  });
}

inline int64_t slab_rows(const m1_ctx* ctx, int batch, int64_t voxels, int per_sm = 4) {
  // ~per_sm blocks per SM over the whole launch
  int64_t slabs = std::max<int64_t>(1, (int64_t)ctx->num_sms * per_sm / std::max(1, batch));
  int64_t rows = cdiv64(voxels, slabs);
  return std::max<int64_t>(rows, 32);
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
  a.mask = d ? d->mask : nullptr;
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
  DISPATCH_T_VW(dtype, C, (inorm_sums_kernel<T, VW><<<grid, TB, 2 * C * sizeof(float), st>>>(
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
  const int64_t rows = slab_rows(ctx, batch, voxels, 8);
  dim3 grid((unsigned)cdiv64(voxels, rows), (unsigned)batch);
  DISPATCH_T_VW(dtype, C, (inorm_act_fwd_kernel<T, VW><<<grid, TB, 0, st>>>(
                             reinterpret_cast<const T*>(x), stats, gamma, beta, voxels, C, slope,
                             reinterpret_cast<T*>(y), rows)));
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
  {
    const int64_t rows8 = slab_rows(ctx, batch, voxels, 8);
    dim3 grid8((unsigned)cdiv64(voxels, rows8), (unsigned)batch);
    DISPATCH_T_VW4(dtype, C, (inorm_bwd_apply_kernel<T, VW><<<grid8, TB, 0, st>>>(
                               reinterpret_cast<const T*>(dy), reinterpret_cast<const T*>(x), stats, gamma, beta, red,
                               voxels, C, slope, 1.f / (float)voxels, reinterpret_cast<T*>(dx), accumulate, rows8)));
  }
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
  DISPATCH_T_VW(dtype, C, (se_squeeze_kernel<T, VW><<<grid, TB, C * sizeof(float), st>>>(
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
  M1_CHECK(dr.mask == nullptr || C % 8 == 0, "m1_se_gate_fwd: the dropout keep-mask needs C %% 8 == 0 (C = %d)", C);
  const int64_t rows = slab_rows(ctx, batch, voxels, 8);
  dim3 grid((unsigned)cdiv64(voxels, rows), (unsigned)batch);
  DISPATCH_T_VW(dtype, C, (se_gate_fwd_kernel<T, VW><<<grid, TB, 0, (cudaStream_t)stream>>>(
                             reinterpret_cast<const T*>(raw3), reinterpret_cast<const T*>(raw4), a, dr, voxels, C,
                             reinterpret_cast<T*>(out), rows)));
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
  {
    const int64_t rows = slab_rows(ctx, batch, voxels, 8);
    dim3 grid((unsigned)cdiv64(voxels, rows), (unsigned)batch);
    DISPATCH_T_VW4(dtype, C, (se_gate_bwd_apply_kernel<T, VW><<<grid, TB, 0, st>>>(
                               reinterpret_cast<const T*>(dout), reinterpret_cast<const T*>(raw3),
                               reinterpret_cast<const T*>(raw4), a, dr, red, voxels, C, 1.f / (float)voxels,
                               reinterpret_cast<T*>(draw3), reinterpret_cast<T*>(draw4), rows)));
  }
  M1_LAUNCH_CHECK(ctx);
  // norm3: dgamma += sum_n A2, dbeta += sum_n (A1 + dpool) ; norm4: dgamma += sum_n B2, dbeta += sum_n B1
  param_grad_kernel<<<(C + 127) / 128, 128, 0, st>>>(red, 5, 1, 0, batch, C, dpool, dgamma3, dbeta3);
  M1_LAUNCH_CHECK(ctx);
  param_grad_kernel<<<(C + 127) / 128, 128, 0, st>>>(red, 5, 3, 2, batch, C, nullptr, dgamma4, dbeta4);
  M1_LAUNCH_CHECK(ctx);
  return 0;
}
