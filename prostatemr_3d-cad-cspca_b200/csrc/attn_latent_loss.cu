// attn_latent_loss.cu — K6 additive attention gate, K7 probabilistic latent heads + analytic KL,
// K8 softmax + focal loss (forward and backward in one pass).
//   K6  GridAttentionBlock3D.call                R:network_blocks.py:106-130
//   K7  mu/log-sigma heads, sample, KL(q||p)     R:networks.py:637-649 (x4), :373-385
//   K8  softmax + Focal.FL / Focal.loss          R:networks.py:388-390,751-755; losses.py:32-49
#include "common.cuh"
#include <algorithm>
#include <cstdlib>

namespace {

constexpr int TB = 256;

struct Grid3 { int d, h, w; };

// ---------------------------------------------------------------------------------------------
// K6 forward: one warp per theta voxel -> psi; then y = up(psi) * x elementwise
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(TB) attn_psi_kernel(const T* __restrict__ theta, const T* __restrict__ phi,
                                                     const float* __restrict__ w_psi,
                                                     const float* __restrict__ b_psi, int batch, Grid3 tg, Grid3 gg,
                                                     int F, float* __restrict__ psi) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (TB / 32);
  const int64_t tvox = (int64_t)tg.d * tg.h * tg.w;
  const int sd = tg.d / gg.d, sh = tg.h / gg.h, sw = tg.w / gg.w;
  for (int64_t v = (int64_t)blockIdx.x * (TB / 32) + (threadIdx.x >> 5); v < (int64_t)batch * tvox; v += warps) {
    const int n = (int)(v / tvox);
    int64_t r = v % tvox;
    const int x = (int)(r % tg.w); r /= tg.w;
    const int y = (int)(r % tg.h);
    const int z = (int)(r / tg.h);
    const int64_t gv = (((int64_t)n * gg.d + min(z / sd, gg.d - 1)) * gg.h + min(y / sh, gg.h - 1)) * gg.w +
                       min(x / sw, gg.w - 1);
    float s = 0.f;
    for (int c = lane; c < F; c += 32) {
      const float f = lrelu(ld_f<T>(theta + v * F + c) + ld_f<T>(phi + gv * F + c), M1_LRELU_SLOPE);
      s = fmaf(f, w_psi[c], s);
    }
    s = warp_sum(s);
    if (lane == 0) psi[v] = 1.f / (1.f + __expf(-(s + b_psi[0])));
  }
}

__device__ __forceinline__ int64_t psi_index(int64_t xv, int n_unused, Grid3 xg, Grid3 tg, int64_t* n_out) {
  int64_t r = xv;
  const int x = (int)(r % xg.w); r /= xg.w;
  const int y = (int)(r % xg.h); r /= xg.h;
  const int z = (int)(r % xg.d); r /= xg.d;
  const int n = (int)r;
  const int sd = xg.d / tg.d, sh = xg.h / tg.h, sw = xg.w / tg.w;
  *n_out = n;
  return (((int64_t)n * tg.d + min(z / sd, tg.d - 1)) * tg.h + min(y / sh, tg.h - 1)) * tg.w + min(x / sw, tg.w - 1);
}

template <typename T>
__global__ void __launch_bounds__(TB) attn_apply_kernel(const T* __restrict__ x, const float* __restrict__ psi,
                                                       Grid3 xg, Grid3 tg, int Cx, T* __restrict__ y,
                                                       __nv_bfloat16* __restrict__ y2, int64_t total) {
  for (int64_t i = blockIdx.x * (int64_t)TB + threadIdx.x; i < total; i += (int64_t)gridDim.x * TB) {
    int64_t n;
    const int64_t pv = psi_index(i / Cx, 0, xg, tg, &n);
    const float v = ld_f<T>(x + i) * psi[pv];
    st_f<T>(y + i, v);
    if (y2) st_f<__nv_bfloat16>(y2 + i, v);
  }
}

// dx (+)= dy * up(psi)
template <typename T>
__global__ void __launch_bounds__(TB) attn_bwd_x_kernel(const T* __restrict__ dy, const float* __restrict__ psi,
                                                       Grid3 xg, Grid3 tg, int Cx, T* __restrict__ dx, int acc,
                                                       int64_t total) {
  for (int64_t i = blockIdx.x * (int64_t)TB + threadIdx.x; i < total; i += (int64_t)gridDim.x * TB) {
    int64_t n;
    const int64_t pv = psi_index(i / Cx, 0, xg, tg, &n);
    float v = ld_f<T>(dy + i) * psi[pv];
    if (acc) v += ld_f<T>(dx + i);
    st_f<T>(dx + i, v);
  }
}

// one warp per theta voxel: dpsi = sum over its x block of <dy, x>; ds = dpsi psi (1-psi);
// dtheta = ds w_psi lrelu'(theta+phi); dphi += dtheta (atomics, fp32); dw_psi/db_psi via smem
template <typename T, typename TG>
__global__ void __launch_bounds__(TB) attn_bwd_psi_kernel(const TG* __restrict__ dy, const T* __restrict__ theta,
                                                         const T* __restrict__ phi, const float* __restrict__ w_psi,
                                                         const float* __restrict__ psi, const T* __restrict__ x,
                                                         int batch, Grid3 tg, Grid3 gg, Grid3 xg, int F, int Cx,
                                                         TG* __restrict__ dtheta, float* __restrict__ dphi,
                                                         float* __restrict__ dw_psi, float* __restrict__ db_psi) {
  extern __shared__ float sdw[];  // F + 1
  for (int i = threadIdx.x; i <= F; i += TB) sdw[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (TB / 32);
  const int64_t tvox = (int64_t)tg.d * tg.h * tg.w;
  const int sd = tg.d / gg.d, sh = tg.h / gg.h, sw = tg.w / gg.w;
  const int ud = xg.d / tg.d, uh = xg.h / tg.h, uw = xg.w / tg.w;
  for (int64_t v = (int64_t)blockIdx.x * (TB / 32) + (threadIdx.x >> 5); v < (int64_t)batch * tvox; v += warps) {
    const int n = (int)(v / tvox);
    int64_t r = v % tvox;
    const int tx = (int)(r % tg.w); r /= tg.w;
    const int ty = (int)(r % tg.h);
    const int tz = (int)(r / tg.h);
    // dpsi over the x voxels that read this psi (nearest up-sampling by (ud,uh,uw))
    float dpsi = 0.f;
    for (int a = 0; a < ud; ++a)
      for (int b = 0; b < uh; ++b)
        for (int c = 0; c < uw; ++c) {
          const int64_t xv = (((int64_t)n * xg.d + tz * ud + a) * xg.h + ty * uh + b) * xg.w + tx * uw + c;
          for (int ch = lane; ch < Cx; ch += 32)
            dpsi = fmaf(ld_f<TG>(dy + xv * Cx + ch), ld_f<T>(x + xv * Cx + ch), dpsi);
        }
    // x voxels beyond tg*u (floor ratio remainder) map to the last psi: handled only when grids divide
    dpsi = warp_sum(dpsi);
    const float p = psi[v];
    const float ds = dpsi * p * (1.f - p);
    const int64_t gv = (((int64_t)n * gg.d + min(tz / sd, gg.d - 1)) * gg.h + min(ty / sh, gg.h - 1)) * gg.w +
                       min(tx / sw, gg.w - 1);
    for (int c = lane; c < F; c += 32) {
      const float pre = ld_f<T>(theta + v * F + c) + ld_f<T>(phi + gv * F + c);
      const float f = lrelu(pre, M1_LRELU_SLOPE);
      const float dth = ds * w_psi[c] * (pre > 0.f ? 1.f : M1_LRELU_SLOPE);
      st_f<TG>(dtheta + v * F + c, dth);
      atomicAdd(dphi + gv * F + c, dth);
      atomicAdd(&sdw[c], ds * f);
    }
    if (lane == 0) atomicAdd(&sdw[F], ds);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < F; i += TB) atomicAdd(dw_psi + i, sdw[i]);
  if (threadIdx.x == 0) atomicAdd(db_psi, sdw[F]);
}

// ---------------------------------------------------------------------------------------------
// K6, vectorised variants (channel counts that are multiples of 8): a group of G = F/8 lanes owns one
// theta voxel, every lane 8 channels (one 16-byte access per tensor); 32/G consecutive voxels share
// a warp, so the phi-gradient atomics of voxels below the same gating voxel are pre-reduced by
// shuffles (the gating grid is 4-16x coarser along w).
// ---------------------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ void load8(const T* p, float (&v)[8]) { ld8v<T>(p, v); }
template <typename T> __device__ __forceinline__ void store8(T* p, const float (&v)[8]) { st8v<T>(p, v); }
__device__ __forceinline__ float group_sum(float v, int G) {   // sum over the G lanes of a voxel group
  for (int o = 1; o < G; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ int64_t coarse_index(int64_t v, Grid3 fine, Grid3 coarse) {
  int64_t r = v;
  const int x = (int)(r % fine.w); r /= fine.w;
  const int y = (int)(r % fine.h); r /= fine.h;
  const int z = (int)(r % fine.d); r /= fine.d;
  const int sd = fine.d / coarse.d, sh = fine.h / coarse.h, sw = fine.w / coarse.w;
  return ((r * coarse.d + min(z / sd, coarse.d - 1)) * coarse.h + min(y / sh, coarse.h - 1)) * coarse.w +
         min(x / sw, coarse.w - 1);
}

template <typename T>
__global__ void __launch_bounds__(TB) attn_psi_vec_kernel(const T* __restrict__ theta, const T* __restrict__ phi,
                                                         const float* __restrict__ w_psi,
                                                         const float* __restrict__ b_psi, int64_t nvox, Grid3 tg,
                                                         Grid3 gg, int F, int G, float* __restrict__ psi) {
  const int lg = threadIdx.x % G;
  const int64_t per_pass = (int64_t)gridDim.x * (TB / G);
  const int64_t rounds = (nvox + per_pass - 1) / per_pass;
  for (int64_t it = 0; it < rounds; ++it) {
    const int64_t v = it * per_pass + (int64_t)blockIdx.x * (TB / G) + threadIdx.x / G;
    float s = 0.f;
    if (v < nvox) {
      const int64_t gv = coarse_index(v, tg, gg);
      for (int c = lg * 8; c < F; c += G * 8) {
        float t8[8], p8[8], w8[8];
        load8<T>(theta + v * F + c, t8);
        load8<T>(phi + gv * F + c, p8);
        load8<float>(w_psi + c, w8);
#pragma unroll
        for (int i = 0; i < 8; ++i) s = fmaf(lrelu(t8[i] + p8[i], M1_LRELU_SLOPE), w8[i], s);
      }
    }
    s = group_sum(s, G);
    if (v < nvox && lg == 0) psi[v] = 1.f / (1.f + __expf(-(s + b_psi[0])));
  }
}

// y = up(psi) * x  (dx mode: out (+)= up(psi) * dy)
template <typename T>
__global__ void __launch_bounds__(TB) attn_scale_vec_kernel(const T* __restrict__ x, const float* __restrict__ psi,
                                                           Grid3 xg, Grid3 tg, int Cx, T* __restrict__ y, int acc,
                                                           int64_t total8, __nv_bfloat16* __restrict__ y2 = nullptr) {
  const int c8 = Cx / 8;
  for (int64_t i = blockIdx.x * (int64_t)TB + threadIdx.x; i < total8; i += (int64_t)gridDim.x * TB) {
    const int64_t xv = i / c8;
    const float p = psi[coarse_index(xv, xg, tg)];
    float v[8];
    load8<T>(x + i * 8, v);
    if (acc) {
      float o[8];
      load8<T>(y + i * 8, o);
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = fmaf(v[k], p, o[k]);
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] *= p;
    }
    store8<T>(y + i * 8, v);
    if (y2) store8<__nv_bfloat16>(y2 + i * 8, v);
  }
}

template <typename T, typename TG>
__global__ void __launch_bounds__(TB) attn_bwd_psi_vec_kernel(
    const TG* __restrict__ dy, const T* __restrict__ theta, const T* __restrict__ phi,
    const float* __restrict__ w_psi, const float* __restrict__ psi, const T* __restrict__ x, int64_t nvox, Grid3 tg,
    Grid3 gg, Grid3 xg, int F, int G, TG* __restrict__ dtheta, float* __restrict__ dphi,
    float* __restrict__ dw_psi, float* __restrict__ db_psi) {
  extern __shared__ float sdw[];  // F + 1
  for (int i = threadIdx.x; i <= F; i += TB) sdw[i] = 0.f;
  __syncthreads();
  const int lg = threadIdx.x % G, lane = threadIdx.x & 31;
  const int ud = xg.d / tg.d, uh = xg.h / tg.h, uw = xg.w / tg.w;
  const int64_t per_pass = (int64_t)gridDim.x * (TB / G);
  const int64_t rounds = (nvox + per_pass - 1) / per_pass;
  for (int64_t it = 0; it < rounds; ++it) {
    const int64_t v = it * per_pass + (int64_t)blockIdx.x * (TB / G) + threadIdx.x / G;
    const bool ok = v < nvox;
    float dpsi = 0.f;
    int64_t gv = -1;
    if (ok) {
      gv = coarse_index(v, tg, gg);
      int64_t r = v;
      const int tx = (int)(r % tg.w); r /= tg.w;
      const int ty = (int)(r % tg.h); r /= tg.h;
      const int tz = (int)(r % tg.d); r /= tg.d;
      for (int a = 0; a < ud; ++a)
        for (int b = 0; b < uh; ++b)
          for (int c = 0; c < uw; ++c) {
            const int64_t xv = ((r * xg.d + tz * ud + a) * xg.h + ty * uh + b) * xg.w + tx * uw + c;
            for (int ch = lg * 8; ch < F; ch += G * 8) {   // Cx == F for every gate of M1
              float d8[8], x8[8];
              load8<TG>(dy + xv * F + ch, d8);
              load8<T>(x + xv * F + ch, x8);
#pragma unroll
              for (int i = 0; i < 8; ++i) dpsi = fmaf(d8[i], x8[i], dpsi);
            }
          }
    }
    dpsi = group_sum(dpsi, G);
    const float p = ok ? psi[v] : 0.f;
    const float ds = dpsi * p * (1.f - p);
    // do all voxel groups of this warp sit below the same gating voxel? then pre-reduce across them
    const int64_t gv0 = __shfl_sync(0xffffffffu, gv, 0);
    const bool same = __all_sync(0xffffffffu, gv == gv0 && ok);
    for (int c = lg * 8; c < F; c += G * 8) {
      float dth[8], dwp[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) { dth[i] = 0.f; dwp[i] = 0.f; }
      if (ok) {
        float t8[8], p8[8], w8[8];
        load8<T>(theta + v * F + c, t8);
        load8<T>(phi + gv * F + c, p8);
        load8<float>(w_psi + c, w8);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float pre = t8[i] + p8[i];
          dth[i] = ds * w8[i] * (pre > 0.f ? 1.f : M1_LRELU_SLOPE);
          dwp[i] = ds * lrelu(pre, M1_LRELU_SLOPE);
        }
        store8<TG>(dtheta + v * F + c, dth);
      }
      // reduce over the voxel groups of the warp (lanes with equal lg)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        for (int o = G; o < 32; o <<= 1) {
          dwp[i] += __shfl_xor_sync(0xffffffffu, dwp[i], o);
          if (same) dth[i] += __shfl_xor_sync(0xffffffffu, dth[i], o);
        }
      }
      if (lane < G) {
#pragma unroll
        for (int i = 0; i < 8; ++i) atomicAdd(&sdw[c + i], dwp[i]);
      }
      if (same) {
        if (lane < G) {
#pragma unroll
          for (int i = 0; i < 8; ++i) atomicAdd(dphi + gv * F + c + i, dth[i]);
        }
      } else if (ok) {
#pragma unroll
        for (int i = 0; i < 8; ++i) atomicAdd(dphi + gv * F + c + i, dth[i]);
      }
    }
    float dsum = (lg == 0 && ok) ? ds : 0.f;
    dsum = warp_sum(dsum);
    if (lane == 0) atomicAdd(&sdw[F], dsum);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < F; i += TB) atomicAdd(dw_psi + i, sdw[i]);
  if (threadIdx.x == 0) atomicAdd(db_psi, sdw[F]);
}

// ---------------------------------------------------------------------------------------------
// K6, fused variants for gates without sub-sampling (att_sub_samp = (1,1,1): theta lives on x's grid - every gate of
// the benchmarked configuration). ONE pass per direction instead of two kernels each:
//   forward   reads theta, x (+ the 128-4096 x smaller phi), writes y (+ twin) and psi
//   backward  reads dy, x, theta (+ dx when accumulating), writes dx and dtheta, adds dphi / dw_psi / db_psi
// grid = (chunks of the H*W plane, batch * D): all index arithmetic is 32-bit and per plane (the old kernels spent
// more instructions on 64-bit div/mod chains than on the data); UNR voxels per thread are loaded before any is used;
// the psi-weight gradient lives in registers for the whole block and is folded once at its end.
// ---------------------------------------------------------------------------------------------
struct AttnGeo {
  int D, H, W, Dg, Hg, Wg, sd, sh, sw;
};
// 8 consecutive elements as loaded (unconverted): the UNR voxels of a thread wait in 4 registers per tensor, not 8
template <typename T> struct Raw8 {
  uint32_t w[sizeof(T) * 2];
  __device__ __forceinline__ void load(const T* p) {
    const uint4 a = *reinterpret_cast<const uint4*>(p);
    w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w;
    if constexpr (sizeof(T) == 4) {
      const uint4 b = *(reinterpret_cast<const uint4*>(p) + 1);
      w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
    }
  }
  __device__ __forceinline__ void get(float (&v)[8]) const {
    if constexpr (sizeof(T) == 4) {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(w[i]);
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) { const float2 f = cvt2<T>(w[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
    }
  }
};
template <int G> struct AttnUnr { static constexpr int v = G <= 8 ? 4 : (G == 16 ? 2 : 1); };

template <typename T, int G>
__global__ void __launch_bounds__(TB, 2) attn_fwd_fused_kernel(const T* __restrict__ theta, const T* __restrict__ phi,
                                                           const float* __restrict__ w_psi,
                                                           const float* __restrict__ b_psi, const T* __restrict__ x,
                                                           AttnGeo g, int nchunks, float* __restrict__ psi,
                                                           T* __restrict__ y, __nv_bfloat16* __restrict__ y2) {
  constexpr int F = 8 * G, VPI = TB / G, UNR = AttnUnr<G>::v;
  const int lg = threadIdx.x % G, vs = threadIdx.x / G;
  const int slab = blockIdx.y, z = slab % g.D, n = slab / g.D;
  const int HW = g.H * g.W;
  const int64_t vbase = (int64_t)slab * HW;
  const int gslab = (n * g.Dg + min(z / g.sd, g.Dg - 1)) * g.Hg;
  float w8[8];
  load8<float>(w_psi + lg * 8, w8);
  const float bias = b_psi[0];
  for (int chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
    const int hw0 = chunk * (VPI * UNR) + vs;
    Raw8<T> rt[UNR], rx[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int hw = hw0 + u * VPI;
      if (hw < HW) {
        rt[u].load(theta + (vbase + hw) * F + lg * 8);
        rx[u].load(x + (vbase + hw) * F + lg * 8);
      }
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int hw = hw0 + u * VPI;
      const bool ok = hw < HW;
      float s = 0.f;
      if (ok) {
        const int yy = hw / g.W, xx = hw - yy * g.W;
        const int64_t gv = (int64_t)(gslab + min(yy / g.sh, g.Hg - 1)) * g.Wg + min(xx / g.sw, g.Wg - 1);
        float p8[8], t8[8];
        load8<T>(phi + gv * F + lg * 8, p8);
        rt[u].get(t8);
#pragma unroll
        for (int i = 0; i < 8; ++i) s = fmaf(lrelu(t8[i] + p8[i], M1_LRELU_SLOPE), w8[i], s);
      }
      s = group_sum(s, G);
      if (ok) {
        const float p = 1.f / (1.f + __expf(-(s + bias)));
        float o[8];
        rx[u].get(o);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] *= p;
        store8<T>(y + (vbase + hw) * F + lg * 8, o);
        if (y2) store8<__nv_bfloat16>(y2 + (vbase + hw) * F + lg * 8, o);
        if (lg == 0) psi[vbase + hw] = p;
      }
    }
  }
}

template <typename T, typename TG, int G>
__global__ void __launch_bounds__(TB, 2) attn_bwd_fused_kernel(const TG* __restrict__ dy, const T* __restrict__ theta,
                                                           const T* __restrict__ phi, const float* __restrict__ w_psi,
                                                           const float* __restrict__ psi, const T* __restrict__ x,
                                                           AttnGeo g, int nchunks, TG* __restrict__ dx, int acc_dx,
                                                           TG* __restrict__ dtheta, float* __restrict__ dphi,
                                                           float* __restrict__ dw_psi, float* __restrict__ db_psi) {
  constexpr int F = 8 * G, VPI = TB / G, UNR = AttnUnr<G>::v;
  __shared__ float sdw[F + 1];
  for (int i = threadIdx.x; i <= F; i += TB) sdw[i] = 0.f;
  __syncthreads();
  const int lg = threadIdx.x % G, vs = threadIdx.x / G, lane = threadIdx.x & 31;
  const int slab = blockIdx.y, z = slab % g.D, n = slab / g.D;
  const int HW = g.H * g.W;
  const int64_t vbase = (int64_t)slab * HW;
  const int gslab = (n * g.Dg + min(z / g.sd, g.Dg - 1)) * g.Hg;
  float w8[8], dwp[8], dbs = 0.f;
  load8<float>(w_psi + lg * 8, w8);
#pragma unroll
  for (int i = 0; i < 8; ++i) dwp[i] = 0.f;
  for (int chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
    const int hw0 = chunk * (VPI * UNR) + vs;
    Raw8<TG> rd[UNR];
    Raw8<T> rx[UNR], rt[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int hw = hw0 + u * VPI;
      if (hw < HW) {
        const int64_t e = (vbase + hw) * F + lg * 8;
        rd[u].load(dy + e);
        rx[u].load(x + e);
        rt[u].load(theta + e);
      }
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int hw = hw0 + u * VPI;
      const bool ok = hw < HW;
      const int64_t e = (vbase + hw) * F + lg * 8;
      float dpsi = 0.f, p = 0.f, d8[8];
      int64_t gv = -1;
      if (ok) {
        float x8[8];
        rd[u].get(d8);
        rx[u].get(x8);
#pragma unroll
        for (int i = 0; i < 8; ++i) dpsi = fmaf(d8[i], x8[i], dpsi);
        p = psi[vbase + hw];
        const int yy = hw / g.W, xx = hw - yy * g.W;
        gv = (int64_t)(gslab + min(yy / g.sh, g.Hg - 1)) * g.Wg + min(xx / g.sw, g.Wg - 1);
      }
      dpsi = group_sum(dpsi, G);
      const float ds = dpsi * p * (1.f - p);
      float dth[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) dth[i] = 0.f;
      if (ok) {
        float o[8], p8[8], t8[8];
        if (acc_dx) load8<TG>(dx + e, o);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = acc_dx ? fmaf(d8[i], p, o[i]) : d8[i] * p;
        store8<TG>(dx + e, o);
        load8<T>(phi + gv * F + lg * 8, p8);
        rt[u].get(t8);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float pre = t8[i] + p8[i];
          dth[i] = ds * w8[i] * (pre > 0.f ? 1.f : M1_LRELU_SLOPE);
          dwp[i] = fmaf(ds, lrelu(pre, M1_LRELU_SLOPE), dwp[i]);
        }
        store8<TG>(dtheta + e, dth);
        if (lg == 0) dbs += ds;
      }
      // phi gradient: the voxel groups of a warp usually sit below ONE gating voxel (the gating grid is 4-16 x
      // coarser along w) - fold them with shuffles, one atomic per channel; otherwise one per voxel and channel
      if constexpr (G < 32) {
        const int64_t gv0 = __shfl_sync(0xffffffffu, gv, 0);
        const bool same = __all_sync(0xffffffffu, gv == gv0);
        if (same) {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            for (int o = G; o < 32; o <<= 1) dth[i] += __shfl_xor_sync(0xffffffffu, dth[i], o);
          if (lane < G && gv >= 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) atomicAdd(dphi + gv * F + lg * 8 + i, dth[i]);
          }
        } else if (ok) {
#pragma unroll
          for (int i = 0; i < 8; ++i) atomicAdd(dphi + gv * F + lg * 8 + i, dth[i]);
        }
      } else if (ok) {
#pragma unroll
        for (int i = 0; i < 8; ++i) atomicAdd(dphi + gv * F + lg * 8 + i, dth[i]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) atomicAdd(&sdw[lg * 8 + i], dwp[i]);
  if (lg == 0) atomicAdd(&sdw[F], dbs);
  __syncthreads();
  for (int i = threadIdx.x; i < F; i += TB) atomicAdd(dw_psi + i, sdw[i]);
  if (threadIdx.x == 0) atomicAdd(db_psi, sdw[F]);
}

// ---------------------------------------------------------------------------------------------
// K7 latent heads
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float clip01(float ls) { return fminf(fmaxf(ls, -0.1f), 0.1f); }
__device__ __forceinline__ float in_clip(float ls) { return (ls >= -0.1f && ls <= 0.1f) ? 1.f : 0.f; }

template <typename T>
__global__ void latent_fwd_kernel(const float* __restrict__ ml, const float* __restrict__ eps, int mode, int L,
                                  int zc, T* __restrict__ z, int64_t rows) {
  const int64_t total = rows * zc;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / zc;
    const int c = (int)(i % zc);
    float v = 0.f;
    if (c < L) {
      const float mu = ml[r * 2 * L + c];
      v = mu;
      if (mode == 0) v = fmaf(__expf(clip01(ml[r * 2 * L + L + c])), eps[r * L + c], mu);
    }
    st_f<T>(z + i, v);
  }
}

template <typename T>
__global__ void latent_bwd_kernel(const T* __restrict__ dz, const float* __restrict__ ml,
                                  const float* __restrict__ eps, int mode, int L, int zc, float* __restrict__ dml,
                                  int64_t rows) {
  const int64_t total = rows * L;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / L;
    const int c = (int)(i % L);
    const float g = ld_f<T>(dz + r * zc + c);
    dml[r * 2 * L + c] += g;
    if (mode == 0) {
      const float ls = ml[r * 2 * L + L + c];
      dml[r * 2 * L + L + c] += g * eps[r * L + c] * __expf(clip01(ls)) * in_clip(ls);
    }
  }
}

__global__ void __launch_bounds__(TB) kl_fwd_kernel(const float* __restrict__ q, const float* __restrict__ p, int L,
                                                   int64_t rows, float scale, float* __restrict__ out, GridSum gs) {
  __shared__ float sm[TB / 32];
  float acc[1] = {0.f};
  const int64_t total = rows * L;
  for (int64_t i = blockIdx.x * (int64_t)TB + threadIdx.x; i < total; i += (int64_t)gridDim.x * TB) {
    const int64_t r = i / L;
    const int c = (int)(i % L);
    const float mq = q[r * 2 * L + c], mp = p[r * 2 * L + c];
    const float lq = clip01(q[r * 2 * L + L + c]), lp = clip01(p[r * 2 * L + L + c]);
    const float dm = mq - mp;
    acc[0] += (lp - lq) + (__expf(2.f * lq) + dm * dm) * 0.5f * __expf(-2.f * lp) - 0.5f;
  }
  block_sum<1, TB>(acc, sm);
  grid_sum_add<TB>(acc[0], scale, out, gs, sm);
}

__global__ void kl_bwd_kernel(const float* __restrict__ q, const float* __restrict__ p, int L, int64_t rows,
                              float scale, float* __restrict__ dq, float* __restrict__ dp) {
  const int64_t total = rows * L;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / L;
    const int c = (int)(i % L);
    const int64_t im = r * 2 * L + c, is = im + L;
    const float mq = q[im], mp = p[im];
    const float rq = q[is], rp = p[is];
    const float lq = clip01(rq), lp = clip01(rp);
    const float dm = mq - mp;
    const float vq = __expf(2.f * lq), ivp = __expf(-2.f * lp);
    const float gm = scale * dm * ivp;
    dq[im] += gm;
    dp[im] -= gm;
    dq[is] += scale * (-1.f + vq * ivp) * in_clip(rq);
    dp[is] += scale * (1.f - (vq + dm * dm) * ivp) * in_clip(rp);
  }
}

// ---------------------------------------------------------------------------------------------
// K8 softmax + focal, one thread per label voxel (nc <= 8 classes)
// ---------------------------------------------------------------------------------------------
constexpr int MAXC = 8;
struct FocalArgs {
  float alpha[MAXC];
  float gamma;
  int nc;
  Grid3 lg, up;
  int batch;
  int prob_c, head_off;
  int from_probs;     // input already holds probabilities (Focal.FL called directly)
  float loss_scale;   // head_weight / batch
  float grad_scale;   // upstream * head_weight / batch
};

template <typename TY>
__global__ void __launch_bounds__(TB) softmax_focal_kernel(const float* __restrict__ logits, const TY* __restrict__ y,
                                                          FocalArgs a, float* __restrict__ prob,
                                                          float* __restrict__ loss_out,
                                                          float* __restrict__ dlogits, GridSum gs) {
  __shared__ float sm[TB / 32];
  const Grid3 xg{a.lg.d * a.up.d, a.lg.h * a.up.h, a.lg.w * a.up.w};
  const int64_t total = (int64_t)a.batch * xg.d * xg.h * xg.w;
  const bool upsampled = a.up.d * a.up.h * a.up.w > 1;
  float acc[1] = {0.f};
  for (int64_t i = blockIdx.x * (int64_t)TB + threadIdx.x; i < total; i += (int64_t)gridDim.x * TB) {
    int64_t r = i;
    const int xw = (int)(r % xg.w); r /= xg.w;
    const int xh = (int)(r % xg.h); r /= xg.h;
    const int xd = (int)(r % xg.d); r /= xg.d;
    const int64_t lv = ((r * a.lg.d + xd / a.up.d) * a.lg.h + xh / a.up.h) * a.lg.w + xw / a.up.w;
    float l[MAXC], p[MAXC], g[MAXC];
    float mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < MAXC; ++c)
      if (c < a.nc) { l[c] = logits[lv * a.nc + c]; mx = fmaxf(mx, l[c]); }
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < MAXC; ++c)
      if (c < a.nc) { p[c] = a.from_probs ? l[c] : expf(l[c] - mx); sum += p[c]; }
    const float inv = a.from_probs ? 1.f : 1.f / sum;
    float psum = 0.f;
#pragma unroll
    for (int c = 0; c < MAXC; ++c)
      if (c < a.nc) {
        p[c] *= inv;
        psum += p[c];
        if (prob) prob[i * a.prob_c + a.head_off + c] = p[c];
      }
    float fl = 0.f, dot = 0.f;
#pragma unroll
    for (int c = 0; c < MAXC; ++c)
      if (c < a.nc) {
        const float yt = y ? ld_f<TY>(y + i * a.nc + c) : 0.f;   // y == NULL: softmax only
        const float qn = p[c] / psum;                                  // y_pred /= sum(y_pred)
        const float q = fminf(fmaxf(qn, 1e-7f), 1.f - 1e-7f);          // clip(eps, 1-eps)
        const float om = 1.f - q;
        const float pw = powf(om, a.gamma);
        const float nl = -logf(q);
        const float wgt = a.alpha[c] * yt * yt;
        fl = fmaf(wgt * pw, nl, fl);
        // d/dq [(1-q)^gamma * (-log q)], zero outside the clip range
        const float inside = (qn >= 1e-7f && qn <= 1.f - 1e-7f) ? 1.f : 0.f;
        const float dpw = a.gamma == 0.f ? 0.f : a.gamma * powf(om, a.gamma - 1.f);
        g[c] = wgt * inside * (-dpw * nl - pw / q);
        dot = fmaf(p[c], g[c], dot);
      }
    acc[0] += fl;
    if (dlogits) {
#pragma unroll
      for (int c = 0; c < MAXC; ++c)
        if (c < a.nc) {
          const float d = a.grad_scale * p[c] * (g[c] - dot);
          if (upsampled) atomicAdd(dlogits + lv * a.nc + c, d);
          else dlogits[lv * a.nc + c] = d;
        }
    }
  }
  block_sum<1, TB>(acc, sm);
  if (loss_out) grid_sum_add<TB>(acc[0], a.loss_scale, loss_out, gs, sm);
}

// ---------------------------------------------------------------------------------------------
// K8 fused with the final logits convolution: thread per voxel, features row in registers
// ---------------------------------------------------------------------------------------------
template <typename T, typename TG, typename TY, int C, int NC>
__global__ void __launch_bounds__(TB) logits_focal_kernel(const T* __restrict__ feat, const float* __restrict__ w,
                                                         const float* __restrict__ bias, const TY* __restrict__ y,
                                                         FocalArgs a, int64_t total, float* __restrict__ prob,
                                                         float* __restrict__ loss_out, TG* __restrict__ dfeat,
                                                         int acc_dfeat, float* __restrict__ dw,
                                                         float* __restrict__ db, GridSum gs) {
  __shared__ float sw[C * NC + NC];
  __shared__ float sdw[C * NC + NC];
  __shared__ float sm[TB / 32];
  for (int i = threadIdx.x; i < C * NC + NC; i += TB) {
    sw[i] = i < C * NC ? w[i] : bias[i - C * NC];
    sdw[i] = 0.f;
  }
  __syncthreads();
  float accw[C * NC + NC];
#pragma unroll
  for (int i = 0; i < C * NC + NC; ++i) accw[i] = 0.f;
  float acc[1] = {0.f};
  for (int64_t i = blockIdx.x * (int64_t)TB + threadIdx.x; i < total; i += (int64_t)gridDim.x * TB) {
    float f[C];
#pragma unroll
    for (int c = 0; c < C; c += 8) {
      float t8[8];
      load8<T>(feat + i * C + c, t8);
#pragma unroll
      for (int k = 0; k < 8; ++k) f[c + k] = t8[k];
    }
    float l[NC], p[NC], g[NC];
    float mx = -INFINITY;
#pragma unroll
    for (int n = 0; n < NC; ++n) {
      float s = sw[C * NC + n];
#pragma unroll
      for (int c = 0; c < C; ++c) s = fmaf(f[c], sw[c * NC + n], s);
      l[n] = s;
      mx = fmaxf(mx, s);
    }
    float sum = 0.f;
#pragma unroll
    for (int n = 0; n < NC; ++n) { p[n] = expf(l[n] - mx); sum += p[n]; }
    const float inv = 1.f / sum;
    float psum = 0.f;
#pragma unroll
    for (int n = 0; n < NC; ++n) {
      p[n] *= inv;
      psum += p[n];
      if (prob) prob[i * a.prob_c + a.head_off + n] = p[n];
    }
    if (y == nullptr) continue;
    float fl = 0.f, dot = 0.f;
#pragma unroll
    for (int n = 0; n < NC; ++n) {
      const float yt = ld_f<TY>(y + i * NC + n);
      const float qn = p[n] / psum;
      const float q = fminf(fmaxf(qn, 1e-7f), 1.f - 1e-7f);
      const float om = 1.f - q;
      const float pw = powf(om, a.gamma);
      const float nl = -logf(q);
      const float wgt = a.alpha[n] * yt * yt;
      fl = fmaf(wgt * pw, nl, fl);
      const float inside = (qn >= 1e-7f && qn <= 1.f - 1e-7f) ? 1.f : 0.f;
      const float dpw = a.gamma == 0.f ? 0.f : a.gamma * powf(om, a.gamma - 1.f);
      g[n] = wgt * inside * (-dpw * nl - pw / q);
      dot = fmaf(p[n], g[n], dot);
    }
    acc[0] += fl;
    if (dfeat == nullptr) continue;
#pragma unroll
    for (int n = 0; n < NC; ++n) {
      g[n] = a.grad_scale * p[n] * (g[n] - dot);          // dL/dlogit_n
      accw[C * NC + n] += g[n];
    }
    float df[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
      float s = 0.f;
#pragma unroll
      for (int n = 0; n < NC; ++n) {
        s = fmaf(g[n], sw[c * NC + n], s);
        accw[c * NC + n] = fmaf(f[c], g[n], accw[c * NC + n]);
      }
      df[c] = s;
    }
#pragma unroll
    for (int c = 0; c < C; c += 8) {
      float t8[8];
      if (acc_dfeat) {
        load8<TG>(dfeat + i * C + c, t8);
#pragma unroll
        for (int k = 0; k < 8; ++k) t8[k] += df[c + k];
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) t8[k] = df[c + k];
      }
      store8<TG>(dfeat + i * C + c, t8);
    }
  }
  if (dw != nullptr) {
#pragma unroll
    for (int i = 0; i < C * NC + NC; ++i) {
      const float v = warp_sum(accw[i]);
      if ((threadIdx.x & 31) == 0) atomicAdd(&sdw[i], v);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < C * NC + NC; i += TB) {
      if (i < C * NC) atomicAdd(dw + i, sdw[i]);
      else atomicAdd(db + (i - C * NC), sdw[i]);
    }
  }
  block_sum<1, TB>(acc, sm);
  if (loss_out) grid_sum_add<TB>(acc[0], a.loss_scale, loss_out, gs, sm);
}

// ---------------------------------------------------------------------------------------------
// cascade (R:networks.py:109-193): backward of logits conv + softmax for a gradient w.r.t. the PROBABILITIES, and the
// decision fusion fused with the focal loss of the joint prediction
// ---------------------------------------------------------------------------------------------
template <typename T, typename TG, int C, int NC>
__global__ void __launch_bounds__(TB) logits_prob_bwd_kernel(const T* __restrict__ feat, const float* __restrict__ w,
                                                            const float* __restrict__ prob, int prob_c, int head_off,
                                                            const float* __restrict__ dprob, int64_t total,
                                                            TG* __restrict__ dfeat, int acc_dfeat,
                                                            float* __restrict__ dw, float* __restrict__ db) {
  __shared__ float sw[C * NC];
  __shared__ float sdw[C * NC + NC];
  for (int i = threadIdx.x; i < C * NC + NC; i += TB) {
    if (i < C * NC) sw[i] = w[i];
    sdw[i] = 0.f;
  }
  __syncthreads();
  float accw[C * NC + NC];
#pragma unroll
  for (int i = 0; i < C * NC + NC; ++i) accw[i] = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)TB + threadIdx.x; i < total; i += (int64_t)gridDim.x * TB) {
    float p[NC], g[NC];
    float dot = 0.f;
#pragma unroll
    for (int n = 0; n < NC; ++n) {
      p[n] = prob[i * prob_c + head_off + n];
      g[n] = dprob[i * NC + n];
      dot = fmaf(p[n], g[n], dot);
    }
#pragma unroll
    for (int n = 0; n < NC; ++n) {
      g[n] = p[n] * (g[n] - dot);                          // dL/dlogit_n
      accw[C * NC + n] += g[n];
    }
    float f[C];
#pragma unroll
    for (int c = 0; c < C; c += 8) {
      float t8[8];
      load8<T>(feat + i * C + c, t8);
#pragma unroll
      for (int k = 0; k < 8; ++k) f[c + k] = t8[k];
    }
#pragma unroll
    for (int c = 0; c < C; c += 8) {
      float t8[8];
      if (acc_dfeat) {
        load8<TG>(dfeat + i * C + c, t8);
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) t8[k] = 0.f;
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
#pragma unroll
        for (int n = 0; n < NC; ++n) {
          t8[k] = fmaf(g[n], sw[(c + k) * NC + n], t8[k]);
          accw[(c + k) * NC + n] = fmaf(f[c + k], g[n], accw[(c + k) * NC + n]);
        }
      }
      store8<TG>(dfeat + i * C + c, t8);
    }
  }
#pragma unroll
  for (int i = 0; i < C * NC + NC; ++i) {
    const float v = warp_sum(accw[i]);
    if ((threadIdx.x & 31) == 0) atomicAdd(&sdw[i], v);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C * NC + NC; i += TB) {
    if (i < C * NC) atomicAdd(dw + i, sdw[i]);
    else atomicAdd(db + (i - C * NC), sdw[i]);
  }
}

struct FusionArgs {
  float alpha[2];
  float gamma;
  int pc1, ch1, pc2, ch2, strategy;
  float loss_scale, grad_scale;
};

// j = fusion(p1, p2); det1 = [1-p1, p1], det2 = [1-j, j]; focal([1-j, j]) and d/dp1, d/dp2
template <typename TY>
__global__ void __launch_bounds__(TB) fusion_focal_kernel(const float* __restrict__ prob1, const float* __restrict__ prob2,
                                                         const TY* __restrict__ y, FusionArgs a, int64_t total,
                                                         float* __restrict__ det1, float* __restrict__ det2,
                                                         float* __restrict__ loss_out, float* __restrict__ dp1,
                                                         float* __restrict__ dp2, GridSum gs) {
  __shared__ float sm[TB / 32];
  float acc[1] = {0.f};
  for (int64_t i = blockIdx.x * (int64_t)TB + threadIdx.x; i < total; i += (int64_t)gridDim.x * TB) {
    const float pa = prob1[i * a.pc1 + a.ch1], pb = prob2[i * a.pc2 + a.ch2];
    float j, dja, djb;
    if (a.strategy == 0) {
      j = pb; dja = 0.f; djb = 1.f;
    } else if (a.strategy == 1) {
      j = 1.f - (1.f - pa) * (1.f - pb); dja = 1.f - pb; djb = 1.f - pa;
    } else {
      const float num = pa * pb + 1e-9f, den = num + (1.f - pa) * (1.f - pb);
      j = num / den;
      dja = (pb * den - num * (2.f * pb - 1.f)) / (den * den);
      djb = (pa * den - num * (2.f * pa - 1.f)) / (den * den);
    }
    if (det1) { det1[2 * i] = 1.f - pa; det1[2 * i + 1] = pa; }
    if (det2) { det2[2 * i] = 1.f - j; det2[2 * i + 1] = j; }
    if (y == nullptr) continue;
    // Focal.FL on y_pred = [1-j, j] (losses.py:32-39): the renormalisation divides by exactly 1
    const float q[2] = {1.f - j, j};
    float fl = 0.f, gq[2];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const float yt = ld_f<TY>(y + i * 2 + c);
      const float qc = fminf(fmaxf(q[c], 1e-7f), 1.f - 1e-7f);
      const float om = 1.f - qc;
      const float pw = powf(om, a.gamma);
      const float nl = -logf(qc);
      const float wgt = a.alpha[c] * yt * yt;
      fl = fmaf(wgt * pw, nl, fl);
      const float inside = (q[c] >= 1e-7f && q[c] <= 1.f - 1e-7f) ? 1.f : 0.f;
      const float dpw = a.gamma == 0.f ? 0.f : a.gamma * powf(om, a.gamma - 1.f);
      gq[c] = wgt * inside * (-dpw * nl - pw / qc);
    }
    acc[0] += fl;
    const float dj = a.grad_scale * (gq[1] - gq[0]);         // q0 = 1 - j, q1 = j
    if (dp1) dp1[i] = dj * dja;
    if (dp2) dp2[i] = dj * djb;
  }
  block_sum<1, TB>(acc, sm);
  if (loss_out) grid_sum_add<TB>(acc[0], a.loss_scale, loss_out, gs, sm);
}

inline unsigned nblocks(const m1_ctx* ctx, int64_t total, int per = TB) {
  return (unsigned)std::max<int64_t>(1, std::min<int64_t>(cdiv64(total, per), (int64_t)ctx->num_sms * 16));
}
inline Grid3 g3(const int32_t* p) { return Grid3{p[0], p[1], p[2]}; }

// ---- fused attention kernels: geometry, grid, lane-group dispatch
inline bool attn_fused_on() {
  static const int on = getenv("M1_ATTN_FUSED") ? atoi(getenv("M1_ATTN_FUSED")) : 1;
  return on != 0;
}
inline AttnGeo attn_geo(const Grid3& t, const Grid3& g) {
  return AttnGeo{t.d, t.h, t.w, g.d, g.h, g.w, t.d / g.d, t.h / g.h, t.w / g.w};
}
inline dim3 attn_fused_grid(const m1_ctx* ctx, const AttnGeo& geo, int batch, int G, int* nchunks) {
  const int unr = G <= 8 ? 4 : (G == 16 ? 2 : 1);
  const int per = (TB / G) * unr;                                 // voxels per block iteration
  *nchunks = (geo.H * geo.W + per - 1) / per;
  const int slabs = batch * geo.D;
  const int gx = std::max(1, std::min(*nchunks, (ctx->num_sms * 16 + slabs - 1) / slabs));
  return dim3((unsigned)gx, (unsigned)slabs);
}
#define M1_ATTN_G(G, LAUNCH)        \
  do {                              \
    switch (G) {                    \
      case 1: LAUNCH(1); break;     \
      case 2: LAUNCH(2); break;     \
      case 4: LAUNCH(4); break;     \
      case 8: LAUNCH(8); break;     \
      case 16: LAUNCH(16); break;   \
      default: LAUNCH(32); break;   \
    }                               \
  } while (0)

}  // namespace

extern "C" int m1_attn_fwd(m1_ctx* ctx, const void* theta, const void* phi, const float* w_psi,
                           const float* b_psi, const void* x, int dtype, int batch, const int32_t* tg,
                           const int32_t* gg, const int32_t* xg, int F, int Cx, float* psi, void* y,
                           void* y_bf16, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const Grid3 T3 = g3(tg), G3 = g3(gg), X3 = g3(xg);
  __nv_bfloat16* y2 = reinterpret_cast<__nv_bfloat16*>(y_bf16);
  M1_CHECK(T3.d % G3.d == 0 && T3.h % G3.h == 0 && T3.w % G3.w == 0 && X3.d % T3.d == 0 && X3.h % T3.h == 0 &&
               X3.w % T3.w == 0,
           "m1_attn_fwd: grids must nest by integer factors");
  const int64_t tv = (int64_t)batch * T3.d * T3.h * T3.w;
  const int64_t total = (int64_t)batch * X3.d * X3.h * X3.w * Cx;
  const int G = F / 8;
  const bool vec = F % 8 == 0 && Cx % 8 == 0 && G >= 1 && G <= 32 && (G & (G - 1)) == 0;
  if (vec && Cx == F && T3.d == X3.d && T3.h == X3.h && T3.w == X3.w && attn_fused_on()) {
    const AttnGeo geo = attn_geo(T3, G3);
    int nchunks;
    const dim3 grid = attn_fused_grid(ctx, geo, batch, G, &nchunks);
#define M1_ATTN_FWD(GG)                                                                                          \
  M1_DISPATCH_T(dtype, T, (attn_fwd_fused_kernel<T, GG><<<grid, TB, 0, st>>>((const T*)theta, (const T*)phi, w_psi, \
                                                                            b_psi, (const T*)x, geo, nchunks, psi, (T*)y, y2)))
    M1_ATTN_G(G, M1_ATTN_FWD);
#undef M1_ATTN_FWD
    M1_LAUNCH_CHECK(ctx);
    return 0;
  }
  if (vec) {
    const unsigned pb = nblocks(ctx, tv, TB / G), sb = nblocks(ctx, total / 8);
    M1_DISPATCH_T(dtype, T, (attn_psi_vec_kernel<T><<<pb, TB, 0, st>>>((const T*)theta, (const T*)phi, w_psi, b_psi, tv,
                                                                      T3, G3, F, G, psi)));
    M1_LAUNCH_CHECK(ctx);
    M1_DISPATCH_T(dtype, T, (attn_scale_vec_kernel<T><<<sb, TB, 0, st>>>((const T*)x, psi, X3, T3, Cx, (T*)y, 0, total / 8, y2)));
    M1_LAUNCH_CHECK(ctx);
    return 0;
  }
  M1_DISPATCH_T(dtype, T, (attn_psi_kernel<T><<<nblocks(ctx, tv, TB / 32), TB, 0, st>>>((const T*)theta, (const T*)phi,
                                                                                       w_psi, b_psi, batch, T3, G3, F, psi)));
  M1_LAUNCH_CHECK(ctx);
  M1_DISPATCH_T(dtype, T, (attn_apply_kernel<T><<<nblocks(ctx, total), TB, 0, st>>>((const T*)x, psi, X3, T3, Cx, (T*)y, y2, total)));
  M1_LAUNCH_CHECK(ctx);
  return 0;
}

/* dtype: VALUE type of theta / phi / x; dy, dx and dtheta are stored as M1_GRAD_DTYPE(dtype) */
extern "C" int m1_attn_bwd(m1_ctx* ctx, const void* dy, const void* theta, const void* phi, const float* w_psi,
                           const float* psi, const void* x, int dtype, int batch, const int32_t* tg,
                           const int32_t* gg, const int32_t* xg, int F, int Cx, void* dx, int acc_dx,
                           void* dtheta, float* dphi, float* dw_psi, float* db_psi, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const Grid3 T3 = g3(tg), G3 = g3(gg), X3 = g3(xg);
  const int64_t tv = (int64_t)batch * T3.d * T3.h * T3.w;
  const int64_t total = (int64_t)batch * X3.d * X3.h * X3.w * Cx;
  const int G = F / 8;
  const bool vec = F % 8 == 0 && Cx == F && G >= 1 && G <= 32 && (G & (G - 1)) == 0;
  if (vec && T3.d == X3.d && T3.h == X3.h && T3.w == X3.w && attn_fused_on()) {
    const AttnGeo geo = attn_geo(T3, G3);
    int nchunks;
    const dim3 grid = attn_fused_grid(ctx, geo, batch, G, &nchunks);
#define M1_ATTN_BWD(GG)                                                                                              \
  M1_DISPATCH_VG(dtype, T, TG, (attn_bwd_fused_kernel<T, TG, GG><<<grid, TB, 0, st>>>(                                \
                                  (const TG*)dy, (const T*)theta, (const T*)phi, w_psi, psi, (const T*)x, geo, nchunks, \
                                  (TG*)dx, acc_dx, (TG*)dtheta, dphi, dw_psi, db_psi)))
    M1_ATTN_G(G, M1_ATTN_BWD);
#undef M1_ATTN_BWD
    M1_LAUNCH_CHECK(ctx);
    return 0;
  }
  if (vec) {
    const unsigned pb = nblocks(ctx, tv, TB / G), sb = nblocks(ctx, total / 8);
    const size_t sm = (F + 1) * sizeof(float);
    M1_DISPATCH_VG(dtype, T, TG, (attn_bwd_psi_vec_kernel<T, TG><<<pb, TB, sm, st>>>(
                                    (const TG*)dy, (const T*)theta, (const T*)phi, w_psi, psi, (const T*)x, tv, T3, G3, X3,
                                    F, G, (TG*)dtheta, dphi, dw_psi, db_psi)));
    M1_LAUNCH_CHECK(ctx);
    M1_DISPATCH_VG(dtype, T, TG, (attn_scale_vec_kernel<TG><<<sb, TB, 0, st>>>((const TG*)dy, psi, X3, T3, Cx, (TG*)dx,
                                                                              acc_dx, total / 8)));
    M1_LAUNCH_CHECK(ctx);
    return 0;
  }
  M1_DISPATCH_VG(dtype, T, TG, (attn_bwd_psi_kernel<T, TG><<<nblocks(ctx, tv, TB / 32), TB, (F + 1) * sizeof(float), st>>>(
                                  (const TG*)dy, (const T*)theta, (const T*)phi, w_psi, psi, (const T*)x, batch, T3, G3, X3,
                                  F, Cx, (TG*)dtheta, dphi, dw_psi, db_psi)));
  M1_LAUNCH_CHECK(ctx);
  M1_DISPATCH_VG(dtype, T, TG, (attn_bwd_x_kernel<TG><<<nblocks(ctx, total), TB, 0, st>>>((const TG*)dy, psi, X3, T3, Cx,
                                                                                         (TG*)dx, acc_dx, total)));
  M1_LAUNCH_CHECK(ctx);
  return 0;
}

extern "C" int m1_latent_fwd(m1_ctx* ctx, const float* ml, const float* eps, int mode, int batch, int64_t voxels,
                             int L, int zdtype, int zc, void* z, void* stream) {
  M1_CHECK(mode == 1 || eps != nullptr, "m1_latent_fwd: sampling mode needs eps");
  M1_CHECK(zc >= L, "m1_latent_fwd: zc < L");
  const int64_t rows = (int64_t)batch * voxels;
  M1_DISPATCH_T(zdtype, T, (latent_fwd_kernel<T><<<nblocks(ctx, rows * zc), TB, 0, (cudaStream_t)stream>>>(
                               ml, eps, mode, L, zc, (T*)z, rows)));
  M1_LAUNCH_CHECK(ctx);
  return 0;
}

/* zdtype: VALUE type of z; dz is stored as M1_GRAD_DTYPE(zdtype) */
extern "C" int m1_latent_bwd(m1_ctx* ctx, const void* dz, const float* ml, const float* eps, int mode, int batch,
                             int64_t voxels, int L, int zdtype, int zc, float* dml, void* stream) {
  const int64_t rows = (int64_t)batch * voxels;
  M1_DISPATCH_VG(zdtype, T, TG, (latent_bwd_kernel<TG><<<nblocks(ctx, rows * L), TB, 0, (cudaStream_t)stream>>>(
                                   (const TG*)dz, ml, eps, mode, L, zc, dml, rows)));
  M1_LAUNCH_CHECK(ctx);
  return 0;
}

extern "C" int m1_kl_fwd(m1_ctx* ctx, const float* ml_q, const float* ml_p, int batch, int64_t voxels, int L,
                         float* kl_out, void* stream) {
  const int64_t rows = (int64_t)batch * voxels;
  kl_fwd_kernel<<<nblocks(ctx, rows * L), TB, 0, (cudaStream_t)stream>>>(ml_q, ml_p, L, rows, 1.f / (float)batch,
                                                                         kl_out, m1_grid_sum(ctx));
  M1_LAUNCH_CHECK(ctx);
  return 0;
}

extern "C" int m1_kl_bwd(m1_ctx* ctx, const float* ml_q, const float* ml_p, int batch, int64_t voxels, int L,
                         float scale, float* dml_q, float* dml_p, void* stream) {
  const int64_t rows = (int64_t)batch * voxels;
  kl_bwd_kernel<<<nblocks(ctx, rows * L), TB, 0, (cudaStream_t)stream>>>(ml_q, ml_p, L, rows,
                                                                         scale / (float)batch, dml_q, dml_p);
  M1_LAUNCH_CHECK(ctx);
  return 0;
}

extern "C" int m1_softmax_focal(m1_ctx* ctx, const void* logits, int ldtype, const void* y_true, int ydtype,
                                const float* alpha, float gamma, int batch, const int32_t* lg, const int32_t* up,
                                int nc, float* prob, int prob_c, int head_off, float head_weight, float* loss_out,
                                void* dlogits, float grad_scale, void* stream) {
  M1_CHECK(ldtype == M1_F32 || ldtype == M1_PROBS,
           "m1_softmax_focal: logits must be fp32 (ldtype M1_PROBS: fp32 probabilities)");
  M1_CHECK(ldtype != M1_PROBS || dlogits == nullptr, "m1_softmax_focal: no gradient in probability mode");
  M1_CHECK(nc >= 1 && nc <= MAXC, "m1_softmax_focal: nc %d out of range", nc);
  FocalArgs a;
  for (int c = 0; c < MAXC; ++c) a.alpha[c] = (alpha && c < nc) ? alpha[c] : 0.f;   // alpha is a HOST array
  a.gamma = gamma;
  a.nc = nc;
  a.lg = g3(lg);
  a.up = g3(up);
  a.batch = batch;
  a.prob_c = prob_c;
  a.head_off = head_off;
  a.from_probs = ldtype == M1_PROBS;
  a.loss_scale = head_weight / (float)batch;
  a.grad_scale = grad_scale * head_weight / (float)batch;
  const int64_t total = (int64_t)batch * a.lg.d * a.up.d * a.lg.h * a.up.h * a.lg.w * a.up.w;
  cudaStream_t st = (cudaStream_t)stream;
  if (y_true == nullptr) {
    // inference: softmax only (labels absent) - reuse the kernel with zero weights
    M1_CHECK(prob != nullptr, "m1_softmax_focal: nothing to do");
  }
  M1_DISPATCH_T(ydtype, TY, (softmax_focal_kernel<TY><<<nblocks(ctx, total), TB, 0, st>>>(
                                (const float*)logits, (const TY*)y_true, a, prob, loss_out, (float*)dlogits,
                                m1_grid_sum(ctx))));
  M1_LAUNCH_CHECK(ctx);
  return 0;
}

template <typename T, typename TG, typename TY, int C, int NC>
static void launch_logits_focal(m1_ctx* ctx, const void* feat, const float* w, const float* bias, const void* y,
                                const FocalArgs& a, int64_t total, float* prob, float* loss_out, void* dfeat,
                                int acc_dfeat, float* dw, float* db, cudaStream_t st) {
  logits_focal_kernel<T, TG, TY, C, NC><<<nblocks(ctx, total), TB, 0, st>>>(
      (const T*)feat, w, bias, (const TY*)y, a, total, prob, loss_out, (TG*)dfeat, acc_dfeat, dw, db, m1_grid_sum(ctx));
}

/* fdtype: VALUE type of feat; dfeat is stored as M1_GRAD_DTYPE(fdtype). y_true: fp32 or bf16. */
extern "C" int m1_logits_softmax_focal(m1_ctx* ctx, const void* feat, int fdtype, const float* w, const float* bias,
                                       const void* y_true, int ydtype, const float* alpha, float gamma, int batch,
                                       int64_t voxels, int C, int nc, float* prob, int prob_c, int head_off,
                                       float head_weight, float* loss_out, void* dfeat, int acc_dfeat, float* dw,
                                       float* db, float grad_scale, void* stream) {
  M1_CHECK(ydtype == M1_F32 || ydtype == M1_BF16, "m1_logits_softmax_focal: labels must be fp32 or bf16");
  FocalArgs a;
  for (int c = 0; c < MAXC; ++c) a.alpha[c] = (alpha && c < nc) ? alpha[c] : 0.f;
  a.gamma = gamma; a.nc = nc; a.batch = batch; a.prob_c = prob_c; a.head_off = head_off;
  a.lg = Grid3{1, 1, 1}; a.up = Grid3{1, 1, 1};
  a.from_probs = 0;
  a.loss_scale = head_weight / (float)batch;
  a.grad_scale = grad_scale * head_weight / (float)batch;
  const int64_t total = (int64_t)batch * voxels;
  cudaStream_t st = (cudaStream_t)stream;
  const bool yb = ydtype == M1_BF16;
#define M1_LF(CC, NN)                                                                                              \
  if (C == CC && nc == NN) {                                                                                       \
    M1_DISPATCH_VG(fdtype, T, TG, {                                                                                \
      if (yb) launch_logits_focal<T, TG, __nv_bfloat16, CC, NN>(ctx, feat, w, bias, y_true, a, total, prob, loss_out, dfeat, acc_dfeat, dw, db, st); \
      else launch_logits_focal<T, TG, float, CC, NN>(ctx, feat, w, bias, y_true, a, total, prob, loss_out, dfeat, acc_dfeat, dw, db, st);            \
    });                                                                                                            \
    M1_LAUNCH_CHECK(ctx);                                                                                          \
    return 0;                                                                                                      \
  }
  M1_LF(32, 2)
  M1_LF(8, 2)
  M1_LF(16, 2)
  M1_LF(32, 3)
#undef M1_LF
  m1_set_error("m1_logits_softmax_focal: no instantiation for C=%d nc=%d", C, nc);
  return 2;
}

template <typename T, typename TG, int C, int NC>
static void launch_logits_prob_bwd(m1_ctx* ctx, const void* feat, const float* w, const float* prob, int prob_c,
                                   int head_off, const float* dprob, int64_t total, void* dfeat, int acc_dfeat,
                                   float* dw, float* db, cudaStream_t st) {
  logits_prob_bwd_kernel<T, TG, C, NC><<<nblocks(ctx, total), TB, 0, st>>>((const T*)feat, w, prob, prob_c, head_off, dprob,
                                                                         total, (TG*)dfeat, acc_dfeat, dw, db);
}

extern "C" int m1_logits_prob_bwd(m1_ctx* ctx, const void* feat, int fdtype, const float* w, const float* prob,
                                  int prob_c, int head_off, const float* dprob, int64_t rows, int C, int nc,
                                  void* dfeat, int acc_dfeat, float* dw, float* db, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
#define M1_LP(CC, NN)                                                                                              \
  if (C == CC && nc == NN) {                                                                                       \
    M1_DISPATCH_VG(fdtype, T, TG, (launch_logits_prob_bwd<T, TG, CC, NN>(ctx, feat, w, prob, prob_c, head_off, dprob, \
                                                                         rows, dfeat, acc_dfeat, dw, db, st)));    \
    M1_LAUNCH_CHECK(ctx);                                                                                          \
    return 0;                                                                                                      \
  }
  M1_LP(32, 2)
  M1_LP(8, 2)
  M1_LP(16, 2)
#undef M1_LP
  m1_set_error("m1_logits_prob_bwd: no instantiation for C=%d nc=%d", C, nc);
  return 2;
}

extern "C" int m1_fusion_focal(m1_ctx* ctx, const float* prob1, int pc1, int ch1, const float* prob2, int pc2, int ch2,
                               int strategy, const void* y_true, int ydtype, const float* alpha, float gamma, int batch,
                               int64_t voxels, float* det1, float* det2, float weight, float* loss_out, float* dp1,
                               float* dp2, float grad_scale, void* stream) {
  M1_CHECK(strategy >= 0 && strategy <= 2, "m1_fusion_focal: unknown strategy %d", strategy);
  M1_CHECK(ydtype == M1_F32 || ydtype == M1_BF16, "m1_fusion_focal: labels must be fp32 or bf16");
  FusionArgs a;
  a.alpha[0] = alpha ? alpha[0] : 0.f;
  a.alpha[1] = alpha ? alpha[1] : 0.f;
  a.gamma = gamma;
  a.pc1 = pc1; a.ch1 = ch1; a.pc2 = pc2; a.ch2 = ch2; a.strategy = strategy;
  a.loss_scale = weight / (float)batch;
  a.grad_scale = grad_scale * weight / (float)batch;
  const int64_t total = (int64_t)batch * voxels;
  cudaStream_t st = (cudaStream_t)stream;
  if (ydtype == M1_BF16)
    fusion_focal_kernel<__nv_bfloat16><<<nblocks(ctx, total), TB, 0, st>>>(prob1, prob2, (const __nv_bfloat16*)y_true, a,
                                                                          total, det1, det2, loss_out, dp1, dp2,
                                                                          m1_grid_sum(ctx));
  else
    fusion_focal_kernel<float><<<nblocks(ctx, total), TB, 0, st>>>(prob1, prob2, (const float*)y_true, a, total, det1, det2,
                                                                  loss_out, dp1, dp2, m1_grid_sum(ctx));
  M1_LAUNCH_CHECK(ctx);
  return 0;
}
