// norm_se.cu — K4/K5: the bandwidth-bound kernels around every convolution of an SE-ResNet block.
//   K4  tfa InstanceNormalization (eps 1e-3, biased variance, affine) + LeakyReLU(0.1)
//       R:networks.py:473,576; R:network_blocks.py:38-44,55,58,104
//   K5  squeeze (GAP), excite (conv6 -> lrelu -> conv7 -> sigmoid), gate * residual -> lrelu
//       -> dropout, R:network_blocks.py:68-78 + R:network_blocks.py:137-143 (tf.nn.dropout)
// All tensors are [batch][voxels][C] (NDHWC flattened); statistics/parameters are fp32. Forward VALUES are
// fp32 / bf16 / fp16 (TV), activation GRADIENTS fp32 / bf16 / bf16 (TG, m1_grad_dtype).
// Vectorised 8 channels per thread (16-byte 16-bit / 2 x 16-byte fp32 accesses). Every per-(sample, channel)
// reduction is DETERMINISTIC: fixed-order shuffles inside a warp, per-warp slots in shared memory summed in
// slot order, per-block partial sums in a scratch buffer summed in block order by the last block to finish
// (ticket counter) - no floating-point atomics, so two runs of the forward pass are bit-identical.
#include "common.cuh"
#include <algorithm>

namespace {

constexpr int TB = 256;

// ---- generic loaders: VW = 8 / 4 channels per thread or 1 (scalar fallback for odd channel counts); value
// arrays always have 8 slots --------------------
template <typename T, int VW>
__device__ __forceinline__ void ldv(const T* p, float (&v)[8]) {
  if constexpr (VW == 8) {
    ld8v<T>(p, v);
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
    st8v<T>(p, v);
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

// Raw (unconverted) vector of VW channels of ES-byte elements: what a load returns before anything depends on
// it. The skeletons below first issue the loads of UNR rows x NT tensors into RawVec registers and only then
// start converting and computing - explicit memory-level parallelism (these kernels are latency bound on
// long-scoreboard stalls otherwise: ~50 KB must be in flight per SM to saturate HBM3e). The element TYPE is
// only needed by unpack<T>: the streamed tensors of one kernel may mix fp16 values and bf16 gradients.
template <int ES, int VW>
struct RawVec {
  static constexpr int kWords = (ES * VW + 3) / 4;
  uint32_t w[kWords];
  __device__ __forceinline__ void load(const void* base, int64_t elem) {
    const char* p = reinterpret_cast<const char*>(base) + elem * ES;
    if constexpr (ES * VW == 32) {
      const uint4 a = *reinterpret_cast<const uint4*>(p), b = *(reinterpret_cast<const uint4*>(p) + 1);
      w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
    } else if constexpr (ES * VW == 16) {
      const uint4 a = *reinterpret_cast<const uint4*>(p);
      w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w;
    } else if constexpr (ES * VW == 8) {
      const uint2 a = *reinterpret_cast<const uint2*>(p);
      w[0] = a.x; w[1] = a.y;
    } else if constexpr (ES == 4) {
      w[0] = *reinterpret_cast<const uint32_t*>(p);
    } else {
      w[0] = *reinterpret_cast<const uint16_t*>(p);
    }
  }
  template <typename T>
  __device__ __forceinline__ void unpack(float (&v)[8]) const {
    static_assert(sizeof(T) == ES, "element type does not match the raw vector");
    if constexpr (ES == 4) {
#pragma unroll
      for (int i = 0; i < VW; ++i) v[i] = __uint_as_float(w[i]);
    } else if constexpr (VW == 1) {
      v[0] = cvt2<T>(w[0]).x;
    } else {
#pragma unroll
      for (int i = 0; i < VW / 2; ++i) {
        const float2 f = cvt2<T>(w[i]);
        v[2 * i] = f.x;
        v[2 * i + 1] = f.y;
      }
    }
  }
};

// Row-parallel skeletons. grid = (slabs, batch); a thread owns ONE channel group (VW channels) and a row
// lane: the per-channel parameters (statistics, affine, gate, reduction results) are loaded and folded ONCE
// into registers by `prep(cbase)`, then the thread streams rows [slab*rows_per_slab, ...) with stride
// `lanes` - one vector access per tensor and row, no index arithmetic, no parameter traffic in the loop.
// Consecutive threads cover consecutive channel groups of a row, then the next row: fully coalesced.
// src[t]: the NT streamed tensors of this sample (ES-byte elements); F gets the raw vectors of one row.
//
// reduce_rows: F(row, cbase, regs, raw[NT], acc[K][8]) accumulates. Deterministic fold (see the file header):
// warp shuffles -> shared-memory slots -> `part` [(sample, slab)][K*C] in the scratch of the context -> the last
// block of the sample (ticket in counter[sample], self-resetting) sums the slabs in order and calls
// FIN(c, totals[K]) once per channel.
struct ReduceScratch {
  float* part;            // [batch * slabs][K * C]
  unsigned int* counter;  // [batch], zero between launches
};

template <int K, int VW, int UNR, int NT, int ES, typename P, typename F, typename FIN>
__device__ __forceinline__ void reduce_rows(const void* const (&src)[NT], int64_t voxels, int C, int64_t rows_per_slab,
                                            float* smem, ReduceScratch rs, P prep, F f, FIN fin) {
  const int CG = C / VW;
  const int cgs = min(CG, TB);
  const int lanes = TB / cgs;
  const int my_lane = threadIdx.x / cgs;
  // lanes of a warp that own the same channel group (cgs < 32: lane ids congruent mod cgs) fold their partial
  // sums with shuffles first (power of two: every thread of the block is active); slot = warp, else slot = lane
  const bool fold = cgs < 32 && (cgs & (cgs - 1)) == 0;
  const int nslots = fold ? TB / 32 : lanes;
  const int slot = fold ? (int)(threadIdx.x >> 5) : my_lane;
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
        RawVec<ES, VW> raw[UNR][NT];
#pragma unroll
        for (int u = 0; u < UNR; ++u)
#pragma unroll
          for (int t = 0; t < NT; ++t) raw[u][t].load(src[t], off + u * step);
#pragma unroll
        for (int u = 0; u < UNR; ++u) f(r + (int64_t)u * lanes, cbase, regs, raw[u], acc);
      }
      for (; r < r1; r += lanes, off += step) {
        RawVec<ES, VW> raw[NT];
#pragma unroll
        for (int t = 0; t < NT; ++t) raw[t].load(src[t], off);
        f(r, cbase, regs, raw, acc);
      }
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
          for (int i = 0; i < VW; ++i) smem[(slot * K + k) * C + cbase + i] = acc[k][i];
      }
    }
  }
  __syncthreads();
  const int n = blockIdx.y, slabs = gridDim.x;
  float* mine = rs.part + ((int64_t)n * slabs + blockIdx.x) * (K * C);
  for (int i = threadIdx.x; i < K * C; i += TB) {
    float t = 0.f;
    for (int sl = 0; sl < nslots; ++sl) t += smem[sl * K * C + i];
    mine[i] = t;
  }
  __threadfence();
  __syncthreads();
  __shared__ unsigned int s_ticket;
  if (threadIdx.x == 0) s_ticket = atomicInc(rs.counter + n, (unsigned)(slabs - 1));   // wraps to 0: self-resetting
  __syncthreads();
  if (s_ticket != (unsigned)(slabs - 1)) return;
  __threadfence();
  const float* all = rs.part + (int64_t)n * slabs * (K * C);
  for (int c = threadIdx.x; c < C; c += TB) {
    float tot[K];
#pragma unroll
    for (int k = 0; k < K; ++k) tot[k] = 0.f;
    for (int sl = 0; sl < slabs; ++sl)
#pragma unroll
      for (int k = 0; k < K; ++k) tot[k] += __ldcg(all + (int64_t)sl * (K * C) + k * C + c);
    fin(c, tot);
  }
}
// shared-memory floats of reduce_rows<K> for C channels of VW-wide groups
inline size_t reduce_smem(int K, int C, int VW) {
  const int CG = C / VW, cgs = std::min(CG, TB), lanes = TB / cgs;
  const bool fold = cgs < 32 && (cgs & (cgs - 1)) == 0;
  return (size_t)(fold ? TB / 32 : lanes) * K * C * sizeof(float);
}

// stream_rows: pure elementwise pass, F(row, element offset, cbase, regs, raw[NT]) computes and stores one row's
// channel group (element offset = row * C + cbase inside the sample)
template <int VW, int UNR, int NT, int ES, typename P, typename F>
__device__ __forceinline__ void stream_rows(const void* const (&src)[NT], int64_t voxels, int C, int64_t rows_per_slab,
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
      RawVec<ES, VW> raw[UNR][NT];
#pragma unroll
      for (int u = 0; u < UNR; ++u)
#pragma unroll
        for (int t = 0; t < NT; ++t) raw[u][t].load(src[t], off + u * step);
#pragma unroll
      for (int u = 0; u < UNR; ++u) f(r + (int64_t)u * lanes, off + u * step, cbase, regs, raw[u]);
    }
    for (; r < r1; r += lanes, off += step) {
      RawVec<ES, VW> raw[NT];
#pragma unroll
      for (int t = 0; t < NT; ++t) raw[t].load(src[t], off);
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
// stats[n][c] = (mean, rstd) over the voxels of sample n; the last block of a sample finalises
template <typename T, int VW>
__global__ void __launch_bounds__(TB) inorm_stats_kernel(const T* __restrict__ x, int64_t voxels, int C,
                                                        int64_t rows_per_slab, float inv_v, float eps,
                                                        ReduceScratch rs, float* __restrict__ stats) {
  extern __shared__ float smem[];
  const int n = blockIdx.y;
  const void* const src[1] = {x + (int64_t)n * voxels * C};
  float* st = stats + (int64_t)n * C * 2;
  reduce_rows<2, VW, 8, 1, sizeof(T)>(
      src, voxels, C, rows_per_slab, smem, rs, [](int) { return 0; },
      [&](int64_t, int, int, const RawVec<sizeof(T), VW> (&raw)[1], float (&acc)[2][8]) {
        float v[8];
        raw[0].template unpack<T>(v);
#pragma unroll
        for (int i = 0; i < VW; ++i) { acc[0][i] += v[i]; acc[1][i] = fmaf(v[i], v[i], acc[1][i]); }
      },
      [&](int c, const float (&t)[2]) {
        const float mean = t[0] * inv_v;
        const float var = fmaxf(t[1] * inv_v - mean * mean, 0.f);
        st[2 * c] = mean;
        st[2 * c + 1] = rsqrtf(var + eps);
      });
}

template <typename T, int VW>
__global__ void __launch_bounds__(TB) inorm_act_fwd_kernel(const T* __restrict__ x, const float* __restrict__ stats,
                                                          const float* __restrict__ gamma,
                                                          const float* __restrict__ beta, int64_t voxels, int C,
                                                          float slope, T* __restrict__ y,
                                                          __nv_bfloat16* __restrict__ y2, int64_t rows_per_slab) {
  const int n = blockIdx.y;
  const void* const src[1] = {x + (int64_t)n * voxels * C};
  T* yb = y + (int64_t)n * voxels * C;
  __nv_bfloat16* y2b = y2 ? y2 + (int64_t)n * voxels * C : nullptr;      // optional bf16 twin of the output
  stream_rows<VW, 8, 1, sizeof(T)>(src, voxels, C, rows_per_slab,
                  [&](int cbase) { return norm_regs<VW>(stats + (int64_t)n * C * 2, gamma, beta, cbase); },
                  [&](int64_t, int64_t off, int, const NormRegs& q, const RawVec<sizeof(T), VW> (&raw)[1]) {
                    float v[8];
                    raw[0].template unpack<T>(v);
#pragma unroll
                    for (int k = 0; k < VW; ++k) v[k] = lrelu(fmaf(v[k], q.a[k], q.b[k]), slope);
                    stv<T, VW>(yb + off, v);
                    if (y2b) stv<__nv_bfloat16, VW>(y2b + off, v);
                  });
}

// ---------------------------------------------------------------------------------------------
// K4 backward: red[n][c] = (sum g, sum g*xhat), g = dy * act'(y)
// ---------------------------------------------------------------------------------------------
template <typename T, typename TG, int VW>
__global__ void __launch_bounds__(TB, 2) inorm_bwd_reduce_kernel(const TG* __restrict__ dy, const T* __restrict__ x,
                                                             const float* __restrict__ stats,
                                                             const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, int64_t voxels,
                                                             int C, float slope, int64_t rows_per_slab,
                                                             ReduceScratch rs, float* __restrict__ red,
                                                             float* __restrict__ dgamma, float* __restrict__ dbeta) {
  static_assert(sizeof(T) == sizeof(TG), "value and gradient storage must have the same width");
  extern __shared__ float smem[];
  const int n = blockIdx.y;
  const void* const src[2] = {x + (int64_t)n * voxels * C, dy + (int64_t)n * voxels * C};
  float* out = red + (int64_t)n * C * 2;
  reduce_rows<2, VW, 4, 2, sizeof(T)>(
      src, voxels, C, rows_per_slab, smem, rs,
      [&](int cbase) { return norm_regs<VW>(stats + (int64_t)n * C * 2, gamma, beta, cbase); },
      [&](int64_t, int, const NormRegs& q, const RawVec<sizeof(T), VW> (&raw)[2], float (&acc)[2][8]) {
        float v[8], d[8];
        raw[0].template unpack<T>(v);
        raw[1].template unpack<TG>(d);
#pragma unroll
        for (int i = 0; i < VW; ++i) {
          const float xh = (v[i] - q.mean[i]) * q.rstd[i];
          const float gg = d[i] * (fmaf(v[i], q.a[i], q.b[i]) > 0.f ? 1.f : slope);
          acc[0][i] += gg;
          acc[1][i] = fmaf(gg, xh, acc[1][i]);
        }
      },
      [&](int c, const float (&t)[2]) {
        out[2 * c] = t[0];
        out[2 * c + 1] = t[1];
        // dbeta += sum g, dgamma += sum g * xhat: one atomic per (sample, channel) - no extra launch
        if (dbeta) atomicAdd(dbeta + c, t[0]);
        if (dgamma) atomicAdd(dgamma + c, t[1]);
      });
}

struct NormBwdRegs {
  NormRegs q;
  float c1[8], c2[8];      // (sum g) / V, (sum g * xhat) / V
};

template <typename T, typename TG, int VW>
__global__ void __launch_bounds__(TB, 2) inorm_bwd_apply_kernel(const TG* __restrict__ dy, const T* __restrict__ x,
                                                            const float* __restrict__ stats,
                                                            const float* __restrict__ gamma,
                                                            const float* __restrict__ beta,
                                                            const float* __restrict__ red, int64_t voxels, int C,
                                                            float slope, float inv_v, TG* __restrict__ dx,
                                                            int accumulate, int64_t rows_per_slab) {
  const int n = blockIdx.y;
  const int64_t base = (int64_t)n * voxels * C;
  const void* const src[2] = {x + base, dy + base};
  TG* dxb = dx + base;
  stream_rows<VW, 4, 2, sizeof(T)>(src, voxels, C, rows_per_slab,
                  [&](int cbase) {
                    NormBwdRegs w;
                    w.q = norm_regs<VW>(stats + (int64_t)n * C * 2, gamma, beta, cbase);
                    const float* rd = red + ((int64_t)n * C + cbase) * 2;
#pragma unroll
                    for (int k = 0; k < VW; ++k) { w.c1[k] = rd[2 * k] * inv_v; w.c2[k] = rd[2 * k + 1] * inv_v; }
                    return w;
                  },
                  [&](int64_t, int64_t off, int, const NormBwdRegs& w, const RawVec<sizeof(T), VW> (&raw)[2]) {
                    float v[8], d[8], o[8];
                    raw[0].template unpack<T>(v);
                    raw[1].template unpack<TG>(d);
                    if (accumulate) ldv<TG, VW>(dxb + off, o);
#pragma unroll
                    for (int k = 0; k < VW; ++k) {
                      const float xh = (v[k] - w.q.mean[k]) * w.q.rstd[k];
                      const float gg = d[k] * (fmaf(v[k], w.q.a[k], w.q.b[k]) > 0.f ? 1.f : slope);
                      const float t = w.q.a[k] * (gg - w.c1[k] - xh * w.c2[k]);
                      o[k] = accumulate ? o[k] + t : t;
                    }
                    stv<TG, VW>(dxb + off, o);
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
// pool = GAP(norm3(raw3)) = a * mean(raw3) + b with a = rstd * gamma, b = beta - mean * a (Q6: == beta up to
// rounding). The mean IS the reduction of the statistics pass over raw3, so the squeeze needs no second read of
// the tensor: one thread per (sample, channel).
__global__ void se_squeeze_kernel(const float* __restrict__ stats3, const float* __restrict__ gamma3,
                                  const float* __restrict__ beta3, int batch, int C, float* __restrict__ pool) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= batch * C) return;
  const int c = i % C;
  const float mean = stats3[2 * i], a = stats3[2 * i + 1] * gamma3[c];
  pool[i] = fmaf(mean, a, beta3[c] - mean * a);
}

// one block per sample; C <= 2048, Cr <= 256
__global__ void __launch_bounds__(TB) se_excite_fwd_kernel(float* __restrict__ pool, const float* __restrict__ w6,
                                                          const float* __restrict__ b6, const float* __restrict__ w7,
                                                          const float* __restrict__ b7, int C, int Cr,
                                                          float* __restrict__ hidden, float* __restrict__ gate,
                                                          const float* __restrict__ stats3,
                                                          const float* __restrict__ gamma3,
                                                          const float* __restrict__ beta3) {
  extern __shared__ float sm[];  // pool[C] | act[Cr]
  float* sp = sm;
  float* sa = sm + C;
  const int n = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += TB) {
    float v;
    if (stats3 != nullptr) {      // squeeze folded in: pool = GAP(norm3(raw3)) from the statistics (se_squeeze_kernel)
      const int64_t i = (int64_t)n * C + c;
      const float mean = stats3[2 * i], a = stats3[2 * i + 1] * gamma3[c];
      v = fmaf(mean, a, beta3[c] - mean * a);
      pool[i] = v;
    } else {
      v = pool[(int64_t)n * C + c];
    }
    sp[c] = v;
  }
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
                                                          float* __restrict__ dw7, float* __restrict__ db7,
                                                          const float* __restrict__ red5, float* __restrict__ dgamma3,
                                                          float* __restrict__ dbeta3, float* __restrict__ dgamma4,
                                                          float* __restrict__ dbeta4) {
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
    if (red5 != nullptr) {
      // parameter gradients of norm3 / norm4 from the reductions of the gate backward (one atomic per sample):
      // norm3: dgamma += A2, dbeta += A1 + dpool ; norm4: dgamma += B2, dbeta += B1
      const float* r = red5 + ((int64_t)n * C + c) * 5;
      atomicAdd(dgamma3 + c, r[1]);
      atomicAdd(dbeta3 + c, r[0] + dp);
      atomicAdd(dgamma4 + c, r[3]);
      atomicAdd(dbeta4 + c, r[2]);
    }
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
__global__ void __launch_bounds__(TB, 2) se_gate_fwd_kernel(const T* __restrict__ raw3, const T* __restrict__ raw4,
                                                        GateArgs a, DropArgs dr, int64_t voxels, int C,
                                                        T* __restrict__ out, __nv_bfloat16* __restrict__ out2,
                                                        int64_t rows_per_slab) {
  const int n = blockIdx.y;
  const int64_t base = (int64_t)n * voxels * C;
  const void* const src[2] = {raw3 + base, raw4 + base};
  stream_rows<VW, 4, 2, sizeof(T)>(src, voxels, C, rows_per_slab, [&](int cbase) { return gate_regs<VW>(a, n, C, cbase); },
                  [&](int64_t, int64_t off, int, const GateRegs& w, const RawVec<sizeof(T), VW> (&raw)[2]) {
                    const int64_t e = base + off;
                    float x3[8], x4[8], f[8], o[8];
                    raw[0].template unpack<T>(x3);
                    raw[1].template unpack<T>(x4);
                    drop_factors<VW>(dr, e, f);
#pragma unroll
                    for (int k = 0; k < VW; ++k) {
                      const float x_ = fmaf(x3[k], w.q3.a[k], w.q3.b[k]);
                      const float res = fmaf(x4[k], w.q4.a[k], w.q4.b[k]);
                      o[k] = lrelu(x_ * w.gt[k] * res, M1_LRELU_SLOPE) * f[k];
                    }
                    stv<T, VW>(out + e, o);
                    if (out2) stv<__nv_bfloat16, VW>(out2 + e, o);
                  });
}

// red[n][c] = { sum dx_, sum dx_*xh3, sum dr, sum dr*xh4, sum dz*x_*r } ; dgate[n][c] = the last one
template <typename T, typename TG, int VW>
__global__ void __launch_bounds__(TB, 2) se_gate_bwd_reduce_kernel(const TG* __restrict__ dout, const T* __restrict__ raw3,
                                                               const T* __restrict__ raw4, GateArgs a, DropArgs dr,
                                                               int64_t voxels, int C, int64_t rows_per_slab,
                                                               ReduceScratch rs, float* __restrict__ red5,
                                                               float* __restrict__ dgate) {
  static_assert(sizeof(T) == sizeof(TG), "value and gradient storage must have the same width");
  extern __shared__ float smem[];
  const int n = blockIdx.y;
  const int64_t base = (int64_t)n * voxels * C;
  const void* const src[3] = {raw3 + base, raw4 + base, dout + base};
  float* out = red5 + (int64_t)n * C * 5;
  float* dg = dgate + (int64_t)n * C;
  reduce_rows<5, VW, 4, 3, sizeof(T)>(
      src, voxels, C, rows_per_slab, smem, rs, [&](int cbase) { return gate_regs<VW>(a, n, C, cbase); },
      [&](int64_t r, int cbase, const GateRegs& w, const RawVec<sizeof(T), VW> (&raw)[3], float (&acc)[5][8]) {
        const int64_t e = base + r * C + cbase;
        float x3[8], x4[8], d[8], f[8];
        raw[0].template unpack<T>(x3);
        raw[1].template unpack<T>(x4);
        raw[2].template unpack<TG>(d);
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
      },
      [&](int c, const float (&t)[5]) {
#pragma unroll
        for (int k = 0; k < 5; ++k) out[5 * c + k] = t[k];
        dg[c] = t[4];
      });
}

struct GateBwdRegs {
  GateRegs w;
  float c[4][8];      // the four reduction results / V
};

template <typename T, typename TG, int VW>
__global__ void __launch_bounds__(TB, 2) se_gate_bwd_apply_kernel(const TG* __restrict__ dout, const T* __restrict__ raw3,
                                                              const T* __restrict__ raw4, GateArgs a, DropArgs dr,
                                                              const float* __restrict__ red5, int64_t voxels, int C,
                                                              float inv_v, TG* __restrict__ draw3,
                                                              TG* __restrict__ draw4, int accumulate,
                                                              int64_t rows_per_slab) {
  const int n = blockIdx.y;
  const int64_t base = (int64_t)n * voxels * C;
  const void* const src[3] = {raw3 + base, raw4 + base, dout + base};
  stream_rows<VW, 4, 3, sizeof(T)>(src, voxels, C, rows_per_slab,
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
                  [&](int64_t, int64_t off, int, const GateBwdRegs& g, const RawVec<sizeof(T), VW> (&raw)[3]) {
                    const GateRegs& w = g.w;
                    const int64_t e = base + off;
                    float x3[8], x4[8], d[8], f[8], o3[8], o4[8];
                    raw[0].template unpack<T>(x3);
                    raw[1].template unpack<T>(x4);
                    raw[2].template unpack<TG>(d);
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
                    if (accumulate) {       // raw3 / raw4 shared by several gate kernels (one trunk, several dropouts)
                      float p3[8], p4[8];
                      ldv<TG, VW>(draw3 + e, p3);
                      ldv<TG, VW>(draw4 + e, p4);
#pragma unroll
                      for (int k = 0; k < VW; ++k) { o3[k] += p3[k]; o4[k] += p4[k]; }
                    }
                    stv<TG, VW>(draw3 + e, o3);
                    stv<TG, VW>(draw4 + e, o4);
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
// value type T (+ its gradient type TG) x vector width
#define DISPATCH_T_VW(dtype, C, ...) M1_DISPATCH_VG(dtype, T, TG, DISPATCH_VW_(C, __VA_ARGS__))
#define DISPATCH_T_VW4(dtype, C, ...) M1_DISPATCH_VG(dtype, T, TG, DISPATCH_VW4_(C, __VA_ARGS__))
inline int vw_of(int C) { return C % 8 == 0 ? 8 : C % 4 == 0 ? 4 : 1; }
inline int vw4_of(int C) { return C % 4 == 0 ? 4 : 1; }

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

// scratch of the deterministic reductions for a (slabs x batch) grid of K*C partial sums per block
int reduce_scratch(m1_ctx* ctx, int64_t slabs, int batch, int K, int C, ReduceScratch* rs) {
  M1_CHECK(batch <= kGridSumCounter, "reduction: batch %d exceeds the ticket counters", batch);
  M1_CHECK((size_t)slabs * batch * K * C * sizeof(float) <= ctx->partial_bytes,
           "reduction: partial sums (%lld blocks x %d) exceed the context scratch", (long long)slabs * batch, K * C);
  rs->part = ctx->partial;
  rs->counter = ctx->counters;
  return 0;
}

}  // namespace

extern "C" int m1_inorm_stats(m1_ctx* ctx, const void* x, int dtype, int batch, int64_t voxels, int C,
                              float eps, float* stats, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t rows = slab_rows(ctx, batch, voxels);
  dim3 grid((unsigned)cdiv64(voxels, rows), (unsigned)batch);
  ReduceScratch rs;
  if (reduce_scratch(ctx, grid.x, batch, 2, C, &rs)) return 1;
  DISPATCH_T_VW(dtype, C, (inorm_stats_kernel<T, VW><<<grid, TB, reduce_smem(2, C, VW), st>>>(
                              reinterpret_cast<const T*>(x), voxels, C, rows, 1.f / (float)voxels, eps, rs, stats)));
  M1_LAUNCH_CHECK(ctx);
  return 0;
}

extern "C" int m1_inorm_act_fwd(m1_ctx* ctx, const void* x, const float* stats, const float* gamma,
                                const float* beta, int dtype, int batch, int64_t voxels, int C, float slope,
                                void* y, void* y_bf16, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t rows = slab_rows(ctx, batch, voxels, 8);
  dim3 grid((unsigned)cdiv64(voxels, rows), (unsigned)batch);
  DISPATCH_T_VW(dtype, C, (inorm_act_fwd_kernel<T, VW><<<grid, TB, 0, st>>>(
                             reinterpret_cast<const T*>(x), stats, gamma, beta, voxels, C, slope,
                             reinterpret_cast<T*>(y), reinterpret_cast<__nv_bfloat16*>(y_bf16), rows)));
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
  const int64_t rows = slab_rows(ctx, batch, voxels);
  dim3 grid((unsigned)cdiv64(voxels, rows), (unsigned)batch);
  ReduceScratch rs;
  if (reduce_scratch(ctx, grid.x, batch, 2, C, &rs)) return 1;
  DISPATCH_T_VW4(dtype, C, (inorm_bwd_reduce_kernel<T, TG, VW><<<grid, TB, reduce_smem(2, C, VW), st>>>(
                              reinterpret_cast<const TG*>(dy), reinterpret_cast<const T*>(x), stats, gamma, beta,
                              voxels, C, slope, rows, rs, red, dgamma, dbeta)));
  M1_LAUNCH_CHECK(ctx);
  {
    const int64_t rows8 = slab_rows(ctx, batch, voxels, 8);
    dim3 grid8((unsigned)cdiv64(voxels, rows8), (unsigned)batch);
    DISPATCH_T_VW4(dtype, C, (inorm_bwd_apply_kernel<T, TG, VW><<<grid8, TB, 0, st>>>(
                               reinterpret_cast<const TG*>(dy), reinterpret_cast<const T*>(x), stats, gamma, beta, red,
                               voxels, C, slope, 1.f / (float)voxels, reinterpret_cast<TG*>(dx), accumulate, rows8)));
  }
  M1_LAUNCH_CHECK(ctx);
  return 0;
}

extern "C" int m1_se_squeeze(m1_ctx* ctx, const void* raw3, const float* stats3, const float* gamma3,
                             const float* beta3, int dtype, int batch, int64_t voxels, int C, float* pool,
                             void* stream) {
  (void)raw3; (void)dtype; (void)voxels;      // the statistics of raw3 already hold its mean (see the kernel)
  const int total = batch * C;
  se_squeeze_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(stats3, gamma3, beta3, batch, C, pool);
  M1_LAUNCH_CHECK(ctx);
  return 0;
}

extern "C" int m1_se_excite_fwd(m1_ctx* ctx, float* pool, const float* w6, const float* b6,
                                const float* w7, const float* b7, int batch, int C, int Cr, float* hidden,
                                float* gate, const float* stats3, const float* gamma3, const float* beta3,
                                void* stream) {
  se_excite_fwd_kernel<<<batch, TB, (C + Cr) * sizeof(float), (cudaStream_t)stream>>>(pool, w6, b6, w7, b7, C, Cr,
                                                                                     hidden, gate, stats3, gamma3, beta3);
  M1_LAUNCH_CHECK(ctx);
  return 0;
}

extern "C" int m1_se_excite_bwd(m1_ctx* ctx, const float* dgate, const float* pool, const float* hidden,
                                const float* gate, const float* w6, const float* w7, int batch, int C, int Cr,
                                float* dpool, float* dw6, float* db6, float* dw7, float* db7, const float* red5,
                                float* dgamma3, float* dbeta3, float* dgamma4, float* dbeta4, void* stream) {
  M1_CHECK(red5 == nullptr || (dgamma3 && dbeta3 && dgamma4 && dbeta4),
           "m1_se_excite_bwd: red5 given without the four norm parameter gradients");
  se_excite_bwd_kernel<<<batch, TB, (2 * C + 2 * Cr) * sizeof(float), (cudaStream_t)stream>>>(
      dgate, pool, hidden, gate, w6, w7, C, Cr, dpool, dw6, db6, dw7, db7, red5, dgamma3, dbeta3, dgamma4, dbeta4);
  M1_LAUNCH_CHECK(ctx);
  return 0;
}

extern "C" int m1_se_gate_fwd(m1_ctx* ctx, const void* raw3, const void* raw4, const float* stats3,
                              const float* stats4, const float* gamma3, const float* beta3, const float* gamma4,
                              const float* beta4, const float* gate, const m1_dropout* drop, int dtype, int batch,
                              int64_t voxels, int C, void* out, void* out_bf16, void* stream) {
  GateArgs a{stats3, stats4, gamma3, beta3, gamma4, beta4, gate};
  DropArgs dr = make_drop(drop);
  M1_CHECK(dr.mask == nullptr || C % 8 == 0, "m1_se_gate_fwd: the dropout keep-mask needs C %% 8 == 0 (C = %d)", C);
  const int64_t rows = slab_rows(ctx, batch, voxels, 8);
  dim3 grid((unsigned)cdiv64(voxels, rows), (unsigned)batch);
  DISPATCH_T_VW(dtype, C, (se_gate_fwd_kernel<T, VW><<<grid, TB, 0, (cudaStream_t)stream>>>(
                             reinterpret_cast<const T*>(raw3), reinterpret_cast<const T*>(raw4), a, dr, voxels, C,
                             reinterpret_cast<T*>(out), reinterpret_cast<__nv_bfloat16*>(out_bf16), rows)));
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
  const int64_t rows = slab_rows(ctx, batch, voxels);
  dim3 grid((unsigned)cdiv64(voxels, rows), (unsigned)batch);
  ReduceScratch rs;
  if (reduce_scratch(ctx, grid.x, batch, 5, C, &rs)) return 1;
  DISPATCH_T_VW4(dtype, C, (se_gate_bwd_reduce_kernel<T, TG, VW><<<grid, TB, reduce_smem(5, C, VW), st>>>(
                              reinterpret_cast<const TG*>(dout), reinterpret_cast<const T*>(raw3),
                              reinterpret_cast<const T*>(raw4), a, dr, voxels, C, rows, rs, red, dgate)));
  M1_LAUNCH_CHECK(ctx);
  return 0;
}

extern "C" int m1_se_gate_bwd_apply(m1_ctx* ctx, const void* dout, const void* raw3, const void* raw4,
                                    const float* stats3, const float* stats4, const float* gamma3,
                                    const float* beta3, const float* gamma4, const float* beta4,
                                    const float* gate, const m1_dropout* drop, const float* red,
                                    const float* dpool, int dtype, int batch, int64_t voxels, int C, void* draw3,
                                    void* draw4, int accumulate, float* dgamma3, float* dbeta3, float* dgamma4,
                                    float* dbeta4, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  GateArgs a{stats3, stats4, gamma3, beta3, gamma4, beta4, gate};
  DropArgs dr = make_drop(drop);
  {
    const int64_t rows = slab_rows(ctx, batch, voxels, 8);
    dim3 grid((unsigned)cdiv64(voxels, rows), (unsigned)batch);
    DISPATCH_T_VW4(dtype, C, (se_gate_bwd_apply_kernel<T, TG, VW><<<grid, TB, 0, st>>>(
                               reinterpret_cast<const TG*>(dout), reinterpret_cast<const T*>(raw3),
                               reinterpret_cast<const T*>(raw4), a, dr, red, voxels, C, 1.f / (float)voxels,
                               reinterpret_cast<TG*>(draw3), reinterpret_cast<TG*>(draw4), accumulate, rows)));
  }
  M1_LAUNCH_CHECK(ctx);
  // norm3: dgamma += sum_n A2, dbeta += sum_n (A1 + dpool) ; norm4: dgamma += sum_n B2, dbeta += sum_n B1
  // (NULL pointers: m1_se_excite_bwd already accumulated them from red5 - two launches fewer per block)
  if (dgamma3 || dbeta3) {
    param_grad_kernel<<<(C + 127) / 128, 128, 0, st>>>(red, 5, 1, 0, batch, C, dpool, dgamma3, dbeta3);
    M1_LAUNCH_CHECK(ctx);
  }
  if (dgamma4 || dbeta4) {
    param_grad_kernel<<<(C + 127) / 128, 128, 0, st>>>(red, 5, 3, 2, batch, C, nullptr, dgamma4, dbeta4);
    M1_LAUNCH_CHECK(ctx);
  }
  return 0;
}
