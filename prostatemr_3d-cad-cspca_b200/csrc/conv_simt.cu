// conv_simt.cu — generic fp32-accumulating gather convolution on the CUDA cores.
//
// This is the engine for (a) precision='fp32' (the 1e-4 parity mode of the north star: tensor
// cores have no fp32 multiply), (b) every launch shape the tcgen05 engine does not take yet
// (strided / transposed / tiny-channel layers) and (c) the weight gradient.  One kernel covers
// Conv3D, Conv3DTranspose and both of their data gradients through the two gather modes of
// m1_conv_desc; it never materialises the channel concatenations of the reference.
//
// Reference call-sites: tf.keras.layers.Conv3D / Conv3DTranspose, R:networks.py:472,496-553,
// R:network_blocks.py:37-46,100-103,275 (TF SAME padding; Conv3DTranspose = exact adjoint).
#include "common.cuh"
#include <algorithm>

namespace {

constexpr int BM = 64;   // output voxels per block
constexpr int BN = 64;   // produced channels per block
constexpr int BK = 16;   // reduced channels per step
constexpr int TH = 128;  // threads

struct SimtParams {
  int mode;
  int batch;
  int Di, Hi, Wi, Do, Ho, Wo;
  int kd, kh, kw, sd, sh, sw, pd, ph, pw;
  int nsrc;
  const void* src[M1_MAX_SRC];
  int src_c[M1_MAX_SRC];
  int nout;
  void* out[M1_MAX_OUT];
  int out_c[M1_MAX_OUT];
  const float* w[M1_MAX_OUT * M1_MAX_SRC];   // [j] or, when w_by_src, [j * nsrc + s]
  int w_by_src;
  const float* bias[M1_MAX_OUT];
  int64_t st[M1_MAX_OUT], sr[M1_MAX_OUT], so[M1_MAX_OUT];
  int accumulate;   // bitmask over outputs
  int n_total;
  int out_start[M1_MAX_OUT + 1];  // prefix sums of out_c
  int64_t out_vox;  // batch * Do*Ho*Wo
};

// produced tensor that owns global column n (n is rewritten to the column inside it)
__device__ __forceinline__ int out_of(const SimtParams& p, int& n) {
  int j = 0;
  while (j + 1 < p.nout && n >= p.out_start[j + 1]) ++j;
  n -= p.out_start[j];
  return j;
}

// input voxel index (within the batch-flattened gathered grid) for output voxel (n,d,h,w) and tap
// (a,b,c); -1 when the tap falls into the SAME padding / between the stride phases
__device__ __forceinline__ int64_t gather_index(const SimtParams& p, int n, int d, int h, int w,
                                                int a, int b, int c) {
  int id, ih, iw;
  if (p.mode == M1_CONV_FWD) {
    id = d * p.sd + a - p.pd;
    ih = h * p.sh + b - p.ph;
    iw = w * p.sw + c - p.pw;
  } else {
    id = d + p.pd - a;
    ih = h + p.ph - b;
    iw = w + p.pw - c;
    if (id < 0 || ih < 0 || iw < 0) return -1;
    if (id % p.sd || ih % p.sh || iw % p.sw) return -1;
    id /= p.sd; ih /= p.sh; iw /= p.sw;
  }
  if (id < 0 || id >= p.Di || ih < 0 || ih >= p.Hi || iw < 0 || iw >= p.Wi) return -1;
  return (((int64_t)n * p.Di + id) * p.Hi + ih) * p.Wi + iw;
}

template <typename T, typename TO>
__global__ void __launch_bounds__(TH) conv_simt_kernel(const __grid_constant__ SimtParams p) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Ws[BK][BN + 4];
  __shared__ int vn[BM], vd[BM], vh[BM], vw[BM];
  __shared__ int64_t vin[BM];

  const int tid = threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int tx = tid % 8, ty = tid / 8;  // 8 channel groups x 16 voxel groups

  if (tid < BM) {
    int64_t m = m0 + tid;
    if (m < p.out_vox) {
      int w = (int)(m % p.Wo); m /= p.Wo;
      int h = (int)(m % p.Ho); m /= p.Ho;
      int d = (int)(m % p.Do); m /= p.Do;
      vn[tid] = (int)m; vd[tid] = d; vh[tid] = h; vw[tid] = w;
    } else {
      vn[tid] = -1;
    }
  }
  __syncthreads();

  float acc[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  const int taps = p.kd * p.kh * p.kw;
  for (int tap = 0; tap < taps; ++tap) {
    const int c = tap % p.kw, b = (tap / p.kw) % p.kh, a = tap / (p.kw * p.kh);
    __syncthreads();
    if (tid < BM) vin[tid] = vn[tid] < 0 ? -1 : gather_index(p, vn[tid], vd[tid], vh[tid], vw[tid], a, b, c);
    __syncthreads();
    // skip taps that touch nothing in this tile (common for transposed gathers)
    int any = 0;
    if (tid < BM) any = vin[tid] >= 0;
    any = __syncthreads_or(any);
    if (!any) continue;

    int r_base = 0;  // reduced-channel offset over the virtual concatenation
    for (int s = 0; s < p.nsrc; ++s) {
      const int C = p.src_c[s];
      const T* src = reinterpret_cast<const T*>(p.src[s]);
      for (int c0 = 0; c0 < C; c0 += BK) {
        // ---- A tile: 64 voxels x 16 channels; thread -> voxel tid/2, channels (tid%2)*8..+8
        {
          const int v = tid >> 1, cb = (tid & 1) * 8;
          const int64_t iv = vin[v];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int ch = c0 + cb + q;
            float x = 0.f;
            if (iv >= 0 && ch < C) x = ld_f<T>(src + iv * C + ch);
            As[cb + q][v] = x;
          }
        }
        // ---- W tile: 16 reduced channels x 64 produced channels
        for (int e = tid; e < BK * BN; e += TH) {
          const int kk = e / BN, nn = e % BN;
          const int ch = c0 + kk;
          int n = n0 + nn;
          float x = 0.f;
          if (ch < C && n < p.n_total) {
            const int j = out_of(p, n);
            x = p.w_by_src ? p.w[j * p.nsrc + s][tap * p.st[s] + (int64_t)ch * p.sr[s] + (int64_t)n * p.so[s]]
                           : p.w[j][tap * p.st[j] + (int64_t)(r_base + ch) * p.sr[j] + (int64_t)n * p.so[j]];
          }
          Ws[kk][nn] = x;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
          float av[4], wv[8];
          const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
          av[0] = a4.x; av[1] = a4.y; av[2] = a4.z; av[3] = a4.w;
          const float4 w0 = *reinterpret_cast<const float4*>(&Ws[kk][tx * 8]);
          const float4 w1 = *reinterpret_cast<const float4*>(&Ws[kk][tx * 8 + 4]);
          wv[0] = w0.x; wv[1] = w0.y; wv[2] = w0.z; wv[3] = w0.w;
          wv[4] = w1.x; wv[5] = w1.y; wv[6] = w1.z; wv[7] = w1.w;
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
        }
        __syncthreads();
      }
      r_base += C;
    }
  }

  // ---- epilogue
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + ty * 4 + i;
    if (m >= p.out_vox) continue;
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
      int n = n0 + tx * 8 + jj;
      if (n >= p.n_total) continue;
      const int j = out_of(p, n);
      float v = acc[i][jj];
      if (p.bias[j]) v += p.bias[j][n];
      TO* dst = reinterpret_cast<TO*>(p.out[j]) + m * p.out_c[j] + n;
      if ((p.accumulate >> j) & 1) v += ld_f<TO>(dst);
      st_f<TO>(dst, v);
    }
  }
}

// ---- few-channel convolution ------------------------------------------------------------------
// Launches with <= 16 gathered or <= 16 produced channels (stem 3-4 -> 32, the f/4 = 8 bottleneck at
// full resolution, 1-3 channel latents, 2-6 channel heads and their data gradients) waste most of the
// 64x64x16 tiles above. Here a thread owns ONE produced voxel x 8 produced channels; the weight
// slice [tap][r][32 channels] of the block's channel tile sits in shared memory, gathered rows are
// read straight from global/L1 - exactly the needed FMAs, coalesced 16-byte stores.
constexpr int TN = 32;   // produced channels per block
constexpr int TV = 32;   // voxels per block (x 4 channel groups of 8 = 128 threads)

template <typename T, typename TO>
__global__ void __launch_bounds__(128) conv_thin_kernel(const __grid_constant__ SimtParams p, int k_total) {
  extern __shared__ float wsm[];   // [tap][r][TN]
  const int tid = threadIdx.x;
  const int n0 = blockIdx.y * TN;
  const int taps = p.kd * p.kh * p.kw;
  for (int e = tid; e < taps * k_total * TN; e += 128) {
    const int nn = e % TN;
    const int r = (e / TN) % k_total;
    const int tap = e / (TN * k_total);
    int n = n0 + nn;
    float x = 0.f;
    if (n < p.n_total) {
      const int j = out_of(p, n);
      if (p.w_by_src) {
        int s = 0, rl = r;
        while (s + 1 < p.nsrc && rl >= p.src_c[s]) { rl -= p.src_c[s]; ++s; }
        x = p.w[j * p.nsrc + s][tap * p.st[s] + (int64_t)rl * p.sr[s] + (int64_t)n * p.so[s]];
      } else {
        x = p.w[j][tap * p.st[j] + (int64_t)r * p.sr[j] + (int64_t)n * p.so[j]];
      }
    }
    wsm[e] = x;
  }
  __syncthreads();
  const int g = tid >> 5;                       // channel group of 8 inside the tile
  const int64_t m = (int64_t)blockIdx.x * TV + (tid & 31);
  if (m >= p.out_vox) return;
  int64_t t = m;
  const int w = (int)(t % p.Wo); t /= p.Wo;
  const int h = (int)(t % p.Ho); t /= p.Ho;
  const int d = (int)(t % p.Do); t /= p.Do;
  const int n_img = (int)t;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  for (int tap = 0; tap < taps; ++tap) {
    const int c = tap % p.kw, b = (tap / p.kw) % p.kh, a = tap / (p.kw * p.kh);
    const int64_t iv = gather_index(p, n_img, d, h, w, a, b, c);
    if (iv < 0) continue;
    int r_base = 0;
    for (int s = 0; s < p.nsrc; ++s) {
      const int C = p.src_c[s];
      const T* row = reinterpret_cast<const T*>(p.src[s]) + iv * C;
      const float* wrow = wsm + ((size_t)tap * k_total + r_base) * TN + g * 8;
      for (int ci = 0; ci < C; ++ci) {
        const float x = ld_f<T>(row + ci);
        const float4 w0 = *reinterpret_cast<const float4*>(wrow + (size_t)ci * TN);
        const float4 w1 = *reinterpret_cast<const float4*>(wrow + (size_t)ci * TN + 4);
        acc[0] = fmaf(x, w0.x, acc[0]); acc[1] = fmaf(x, w0.y, acc[1]);
        acc[2] = fmaf(x, w0.z, acc[2]); acc[3] = fmaf(x, w0.w, acc[3]);
        acc[4] = fmaf(x, w1.x, acc[4]); acc[5] = fmaf(x, w1.y, acc[5]);
        acc[6] = fmaf(x, w1.z, acc[6]); acc[7] = fmaf(x, w1.w, acc[7]);
      }
      r_base += C;
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int n = n0 + g * 8 + i;
    if (n >= p.n_total) continue;
    const int j = out_of(p, n);
    float v = acc[i];
    if (p.bias[j]) v += p.bias[j][n];
    TO* dst = reinterpret_cast<TO*>(p.out[j]) + m * p.out_c[j] + n;
    if ((p.accumulate >> j) & 1) v += ld_f<TO>(dst);
    st_f<TO>(dst, v);
  }
}

// ---- 1x1x1 heads ---------------------------------------------------------------------------------
// The mu/log-sigma heads (R:networks.py:637-641: 128/256/512 -> 2/4/6 channels, fp32 out) and their two
// gradients are pure streaming: one read of the feature tensor. Lanes own 8-channel chunks (16-byte
// loads), head weights live in registers, partial dot products meet through warp shuffles.
//   pw_head_fwd   : out[v][n]  = b[n] + sum_r x[v][r] W[r][n]                 (N <= 8)
//   pw_head_dgrad : dx[v][r] (+)= sum_n dy[v][n] W[r][n]                      (N <= 8 gathered)
//   pw_head_wgrad : dW[r][n]  += sum_v x[v][r] dy[v][n]
template <typename T> __device__ __forceinline__ void ld8(const T* p, float (&v)[8]) { ld8v<T>(p, v); }
template <typename T> __device__ __forceinline__ void st8(T* p, const float (&v)[8]) { st8v<T>(p, v); }

struct HeadParams {
  const void* x;      // [vox][C] features (fwd, wgrad)
  const float* dy;    // [vox][N] fp32 (dgrad, wgrad)
  void* out;          // fwd: float [vox][N]; dgrad: T [vox][C]
  const float* w;     // W[r * sr + n * so]
  const float* bias;
  float* dw;
  int64_t sr, so;
  int64_t vox;
  int C, N, accumulate;
};

// lanes per voxel LPV = min(32, C/8); a lane owns chunks lane%LPV + i*32 (i < CPL)
template <typename T, int NP, int CPL>
__global__ void __launch_bounds__(256) pw_head_fwd_kernel(const HeadParams p) {
  const int chunks = p.C / 8;
  const int lpv = chunks < 32 ? chunks : 32;
  const int lane = threadIdx.x & 31, sub = lane % lpv, vpw = 32 / lpv;   // voxels per warp iteration
  float w[CPL][8][NP];
#pragma unroll
  for (int i = 0; i < CPL; ++i)
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int n = 0; n < NP; ++n) {
        const int ch = (sub + i * 32) * 8 + r;
        w[i][r][n] = (n < p.N && ch < p.C) ? p.w[(int64_t)ch * p.sr + (int64_t)n * p.so] : 0.f;
      }
  float b[NP];
#pragma unroll
  for (int n = 0; n < NP; ++n) b[n] = (p.bias && n < p.N) ? p.bias[n] : 0.f;
  const T* x = reinterpret_cast<const T*>(p.x);
  float* out = reinterpret_cast<float*>(p.out);
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5), wid = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  for (int64_t v0 = wid * vpw; v0 < p.vox; v0 += warps * vpw) {
    const int64_t v = v0 + lane / lpv;
    float acc[NP];
#pragma unroll
    for (int n = 0; n < NP; ++n) acc[n] = 0.f;
    if (v < p.vox) {
#pragma unroll
      for (int i = 0; i < CPL; ++i) {
        const int c0 = (sub + i * 32) * 8;
        if (c0 < p.C) {
          float xv[8];
          ld8<T>(x + v * p.C + c0, xv);
#pragma unroll
          for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int n = 0; n < NP; ++n) acc[n] = fmaf(xv[r], w[i][r][n], acc[n]);
        }
      }
    }
    for (int o = lpv >> 1; o > 0; o >>= 1)
#pragma unroll
      for (int n = 0; n < NP; ++n) acc[n] += __shfl_xor_sync(0xffffffffu, acc[n], o);
    if (sub == 0 && v < p.vox) {
#pragma unroll
      for (int n = 0; n < NP; ++n)
        if (n < p.N) {
          float r = acc[n] + b[n];
          if (p.accumulate) r += out[v * p.N + n];
          out[v * p.N + n] = r;
        }
    }
  }
}

// thread = (voxel lane, 8-channel chunk): blockDim.x % (C/8) == 0
template <typename T, int NP>
__global__ void __launch_bounds__(256) pw_head_dgrad_kernel(const HeadParams p) {
  const int chunks = p.C / 8;
  const int chunk = threadIdx.x % chunks, vl = threadIdx.x / chunks, vpb = blockDim.x / chunks;
  float w[8][NP];
#pragma unroll
  for (int r = 0; r < 8; ++r)
#pragma unroll
    for (int n = 0; n < NP; ++n) w[r][n] = n < p.N ? p.w[(int64_t)(chunk * 8 + r) * p.sr + (int64_t)n * p.so] : 0.f;
  T* dx = reinterpret_cast<T*>(p.out);
  for (int64_t v = (int64_t)blockIdx.x * vpb + vl; v < p.vox; v += (int64_t)gridDim.x * vpb) {
    float g[NP];
#pragma unroll
    for (int n = 0; n < NP; ++n) g[n] = n < p.N ? p.dy[v * p.N + n] : 0.f;
    float o[8];
    T* dst = dx + v * p.C + chunk * 8;
    if (p.accumulate) {
      ld8<T>(dst, o);
    } else {
#pragma unroll
      for (int r = 0; r < 8; ++r) o[r] = 0.f;
    }
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int n = 0; n < NP; ++n) o[r] = fmaf(g[n], w[r][n], o[r]);
    st8<T>(dst, o);
  }
}

// thread = (voxel lane, 8-channel chunk); 8 x N accumulators, block reduction through shared memory
template <typename T, int NP>
__global__ void __launch_bounds__(256) pw_head_wgrad_kernel(const HeadParams p) {
  extern __shared__ float red[];   // [C][NP]
  const int chunks = p.C / 8;
  const int chunk = threadIdx.x % chunks, vl = threadIdx.x / chunks, vpb = blockDim.x / chunks;
  for (int i = threadIdx.x; i < p.C * NP; i += blockDim.x) red[i] = 0.f;
  __syncthreads();
  float acc[8][NP];
#pragma unroll
  for (int r = 0; r < 8; ++r)
#pragma unroll
    for (int n = 0; n < NP; ++n) acc[r][n] = 0.f;
  const T* x = reinterpret_cast<const T*>(p.x);
  for (int64_t v = (int64_t)blockIdx.x * vpb + vl; v < p.vox; v += (int64_t)gridDim.x * vpb) {
    float g[NP], xv[8];
#pragma unroll
    for (int n = 0; n < NP; ++n) g[n] = n < p.N ? p.dy[v * p.N + n] : 0.f;
    ld8<T>(x + v * p.C + chunk * 8, xv);
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int n = 0; n < NP; ++n) acc[r][n] = fmaf(xv[r], g[n], acc[r][n]);
  }
#pragma unroll
  for (int r = 0; r < 8; ++r)
#pragma unroll
    for (int n = 0; n < NP; ++n) atomicAdd(&red[(chunk * 8 + r) * NP + n], acc[r][n]);
  __syncthreads();
  for (int i = threadIdx.x; i < p.C * NP; i += blockDim.x) {
    const int r = i / NP, n = i % NP;
    if (n < p.N && red[i] != 0.f) atomicAdd(p.dw + (int64_t)r * p.sr + (int64_t)n * p.so, red[i]);
  }
}

// chunks = C/8 must divide 256 (dgrad/wgrad thread mapping) and be a power of two <= 32 or a multiple of 32 (fwd)
inline bool head_shape_ok(int C, int N) {
  if (N < 1 || N > 8 || C % 8) return false;
  const int chunks = C / 8;
  return chunks == 4 || chunks == 8 || chunks == 16 || chunks == 32 || chunks == 64;
}
inline bool pointwise_s1(const m1_conv_desc* d) {
  for (int i = 0; i < 3; ++i)
    if (d->kernel[i] != 1 || d->stride[i] != 1) return false;
  return d->nsrc == 1 && d->nout == 1 && !d->w_by_src;
}

template <typename T>
int launch_head(m1_ctx* ctx, int which, const HeadParams& p, cudaStream_t st) {
  const int chunks = p.C / 8;
  const int np = (p.N + 1) & ~1;   // instantiated widths: 2, 4, 6, 8
  const int64_t vpb = which == 0 ? (chunks >= 32 ? 8 : 8 * (32 / chunks)) : 256 / chunks;
  const int threads = 256;
  int64_t blocks = std::min<int64_t>(cdiv64(p.vox, vpb), (int64_t)ctx->num_sms * (which == 2 ? 4 : 16));
  if (blocks < 1) blocks = 1;
#define M1_HEAD_CASE(NP)                                                                                        \
  case NP:                                                                                                      \
    if (which == 0) {                                                                                           \
      if (chunks > 32) pw_head_fwd_kernel<T, NP, 2><<<(unsigned)blocks, 256, 0, st>>>(p);                       \
      else pw_head_fwd_kernel<T, NP, 1><<<(unsigned)blocks, 256, 0, st>>>(p);                                   \
    } else if (which == 1) {                                                                                    \
      pw_head_dgrad_kernel<T, NP><<<(unsigned)blocks, threads, 0, st>>>(p);                                     \
    } else {                                                                                                    \
      pw_head_wgrad_kernel<T, NP><<<(unsigned)blocks, threads, (size_t)p.C * NP * sizeof(float), st>>>(p);      \
    }                                                                                                           \
    break;
  switch (np) {
    M1_HEAD_CASE(2)
    M1_HEAD_CASE(4)
    M1_HEAD_CASE(6)
    M1_HEAD_CASE(8)
    default: m1_set_error("head kernel: unsupported width %d", p.N); return 1;
  }
#undef M1_HEAD_CASE
  M1_LAUNCH_CHECK(ctx);
  return 0;
}

// ---- weight gradient ----------------------------------------------------------------------
// dW[tap, r, n] += sum_o G(o, tap)[r] * dY[o, n];  block = (64 r x 64 n) tile of one tap over a
// slab of output voxels, fp32 atomics into dW.
constexpr int WV = 16;  // voxels per step

struct WgradParams {
  SimtParams g;          // geometry + gathered tensors (out/w unused)
  const void* dout;      // [out_vox][Cn]
  int Cn;
  int src_index;         // which gathered tensor this launch differentiates
  int r_base;            // its offset in the concatenation
  float* dw;
  int64_t st, sr, so;
  int64_t vox_per_block;
};

template <typename T, typename TO>
__global__ void __launch_bounds__(TH) wgrad_simt_kernel(const __grid_constant__ WgradParams q) {
  const SimtParams& p = q.g;
  __shared__ float Gs[WV][BM + 4];
  __shared__ float Ys[WV][BN + 4];
  __shared__ int64_t vin[WV];

  const int tid = threadIdx.x;
  const int C = p.src_c[q.src_index];
  const T* src = reinterpret_cast<const T*>(p.src[q.src_index]);
  const TO* dy = reinterpret_cast<const TO*>(q.dout);
  const int r_tiles = (C + BM - 1) / BM;
  const int r0 = (blockIdx.x % r_tiles) * BM;
  const int n0 = (blockIdx.x / r_tiles) * BN;
  const int tap = blockIdx.y;
  const int c = tap % p.kw, b = (tap / p.kw) % p.kh, a = tap / (p.kw * p.kh);
  const int64_t v_begin = (int64_t)blockIdx.z * q.vox_per_block;
  const int64_t v_end = min(v_begin + q.vox_per_block, p.out_vox);
  const int tx = tid % 8, ty = tid / 8;

  float acc[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  for (int64_t v0 = v_begin; v0 < v_end; v0 += WV) {
    if (tid < WV) {
      int64_t m = v0 + tid;
      int64_t iv = -1;
      if (m < v_end) {
        int w = (int)(m % p.Wo); m /= p.Wo;
        int h = (int)(m % p.Ho); m /= p.Ho;
        int d = (int)(m % p.Do); m /= p.Do;
        iv = gather_index(p, (int)m, d, h, w, a, b, c);
      }
      vin[tid] = iv;
    }
    __syncthreads();
    for (int e = tid; e < WV * BM; e += TH) {
      const int vv = e / BM, rr = e % BM;
      const int64_t iv = vin[vv];
      float x = 0.f;
      if (iv >= 0 && r0 + rr < C) x = ld_f<T>(src + iv * C + r0 + rr);
      Gs[vv][rr] = x;
    }
    for (int e = tid; e < WV * BN; e += TH) {
      const int vv = e / BN, nn = e % BN;
      const int64_t m = v0 + vv;
      float x = 0.f;
      if (m < v_end && n0 + nn < q.Cn && vin[vv] >= 0) x = ld_f<TO>(dy + m * q.Cn + n0 + nn);
      Ys[vv][nn] = x;
    }
    __syncthreads();
#pragma unroll
    for (int vv = 0; vv < WV; ++vv) {
      float gv[4], yv[8];
      const float4 g4 = *reinterpret_cast<const float4*>(&Gs[vv][ty * 4]);
      gv[0] = g4.x; gv[1] = g4.y; gv[2] = g4.z; gv[3] = g4.w;
      const float4 y0 = *reinterpret_cast<const float4*>(&Ys[vv][tx * 8]);
      const float4 y1 = *reinterpret_cast<const float4*>(&Ys[vv][tx * 8 + 4]);
      yv[0] = y0.x; yv[1] = y0.y; yv[2] = y0.z; yv[3] = y0.w;
      yv[4] = y1.x; yv[5] = y1.y; yv[6] = y1.z; yv[7] = y1.w;
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(gv[i], yv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = r0 + ty * 4 + i;
    if (r >= C) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = n0 + tx * 8 + j;
      if (n >= q.Cn) continue;
      if (acc[i][j] != 0.f)
        atomicAdd(q.dw + tap * q.st + (int64_t)(q.r_base + r) * q.sr + (int64_t)n * q.so, acc[i][j]);
    }
  }
}

// ---- weight gradient, few-channel layers ------------------------------------------------------
// Layers such as the stem (3-4 -> 32), the f/4 = 8 bottlenecks at full resolution, the 1-3 channel
// latents and the 2-6 channel heads have R*N <= 4096 outputs per tap but millions of voxels: the
// 64x64-tile kernel above would spend >90 % of its FMAs on zero padding. Here a block owns one tap
// and a voxel slab, stages VCH voxels of both operands in shared memory and every thread keeps
// ceil(R*N / blockDim) accumulators in registers - exactly the needed MACs.
// Thread = (4 reduced rows x 8 produced channels) micro-tile x one voxel lane: per staged voxel it
// reads 4 + 8 values from shared memory for 32 FMAs; the voxel lanes of a block split the slab.
constexpr int SW_RB = 4, SW_NB = 8;

template <typename T, typename TO>
__global__ void __launch_bounds__(256) wgrad_small_kernel(const __grid_constant__ WgradParams q, int vch) {
  const SimtParams& p = q.g;
  extern __shared__ float sm[];
  const int R = p.src_c[q.src_index], N = q.Cn;
  const int Rp = (R + SW_RB - 1) / SW_RB * SW_RB, Np = (N + SW_NB - 1) / SW_NB * SW_NB;
  float* xs = sm;                      // [vch][Rp]
  float* ys = sm + (size_t)vch * Rp;   // [vch][Np]
  __shared__ int64_t vin[64];
  const T* src = reinterpret_cast<const T*>(p.src[q.src_index]);
  const TO* dy = reinterpret_cast<const TO*>(q.dout);
  const int tap = blockIdx.y;
  const int c = tap % p.kw, b = (tap / p.kw) % p.kh, a = tap / (p.kw * p.kh);
  const int64_t v_begin = (int64_t)blockIdx.x * q.vox_per_block;
  const int64_t v_end = min(v_begin + q.vox_per_block, p.out_vox);
  const int mt = (Rp / SW_RB) * (Np / SW_NB);      // micro-tiles
  const int lanes = blockDim.x / mt;               // voxel lanes (>= 1 by construction)
  const int my_mt = threadIdx.x % mt, my_lane = threadIdx.x / mt;
  const int r0 = (my_mt / (Np / SW_NB)) * SW_RB, n0 = (my_mt % (Np / SW_NB)) * SW_NB;
  float acc[SW_RB][SW_NB];
#pragma unroll
  for (int i = 0; i < SW_RB; ++i)
#pragma unroll
    for (int j = 0; j < SW_NB; ++j) acc[i][j] = 0.f;

  for (int64_t v0 = v_begin; v0 < v_end; v0 += vch) {
    __syncthreads();
    if (threadIdx.x < vch) {
      int64_t m = v0 + threadIdx.x;
      int64_t iv = -1;
      if (m < v_end) {
        int w = (int)(m % p.Wo); m /= p.Wo;
        int h = (int)(m % p.Ho); m /= p.Ho;
        int d = (int)(m % p.Do); m /= p.Do;
        iv = gather_index(p, (int)m, d, h, w, a, b, c);
      }
      vin[threadIdx.x] = iv;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < vch * Rp; e += blockDim.x) {
      const int vv = e / Rp, rr = e % Rp;
      const int64_t iv = vin[vv];
      xs[e] = (iv >= 0 && rr < R) ? ld_f<T>(src + iv * R + rr) : 0.f;
    }
    for (int e = threadIdx.x; e < vch * Np; e += blockDim.x) {
      const int vv = e / Np, nn = e % Np;
      ys[e] = (vin[vv] >= 0 && nn < N) ? ld_f<TO>(dy + (v0 + vv) * N + nn) : 0.f;
    }
    __syncthreads();
    if (my_lane < lanes) {
      for (int vv = my_lane; vv < vch; vv += lanes) {
        const float4 x4 = *reinterpret_cast<const float4*>(xs + vv * Rp + r0);
        const float4 y0 = *reinterpret_cast<const float4*>(ys + vv * Np + n0);
        const float4 y1 = *reinterpret_cast<const float4*>(ys + vv * Np + n0 + 4);
        const float xv[4] = {x4.x, x4.y, x4.z, x4.w};
        const float yv[8] = {y0.x, y0.y, y0.z, y0.w, y1.x, y1.y, y1.z, y1.w};
#pragma unroll
        for (int i = 0; i < SW_RB; ++i)
#pragma unroll
          for (int j = 0; j < SW_NB; ++j) acc[i][j] = fmaf(xv[i], yv[j], acc[i][j]);
      }
    }
  }
  if (my_lane < lanes) {
#pragma unroll
    for (int i = 0; i < SW_RB; ++i)
#pragma unroll
      for (int j = 0; j < SW_NB; ++j) {
        const int rr = r0 + i, nn = n0 + j;
        if (rr < R && nn < N && acc[i][j] != 0.f)
          atomicAdd(q.dw + tap * q.st + (int64_t)(q.r_base + rr) * q.sr + (int64_t)nn * q.so, acc[i][j]);
      }
  }
}

// column sums: dbias[n] += sum_rows x[row][n]
template <typename T>
__global__ void colsum_kernel(const T* __restrict__ x, int64_t rows, int C, int64_t rows_per_block,
                              float* __restrict__ out) {
  extern __shared__ float sacc[];
  for (int i = threadIdx.x; i < C; i += blockDim.x) sacc[i] = 0.f;
  __syncthreads();
  const int64_t r_begin = (int64_t)blockIdx.x * rows_per_block;
  const int64_t r_end = min(r_begin + rows_per_block, rows);
  if (C <= (int)blockDim.x && blockDim.x % C == 0) {
    const int ch = threadIdx.x % C;
    const int lanes = blockDim.x / C;
    float a = 0.f;
    for (int64_t r = r_begin + threadIdx.x / C; r < r_end; r += lanes) a += ld_f<T>(x + r * C + ch);
    atomicAdd(&sacc[ch], a);
  } else {
    for (int ch = threadIdx.x; ch < C; ch += blockDim.x) {
      float a = 0.f;
      for (int64_t r = r_begin; r < r_end; ++r) a += ld_f<T>(x + r * C + ch);
      sacc[ch] += a;  // each channel owned by exactly one thread on this path
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) atomicAdd(out + i, sacc[i]);
}

int fill_params(const m1_conv_desc* d, SimtParams* p) {
  memset(p, 0, sizeof(*p));
  p->mode = d->mode;
  p->batch = d->batch;
  p->Di = d->in_dhw[0]; p->Hi = d->in_dhw[1]; p->Wi = d->in_dhw[2];
  p->Do = d->out_dhw[0]; p->Ho = d->out_dhw[1]; p->Wo = d->out_dhw[2];
  p->kd = d->kernel[0]; p->kh = d->kernel[1]; p->kw = d->kernel[2];
  p->sd = d->stride[0]; p->sh = d->stride[1]; p->sw = d->stride[2];
  p->pd = d->pad[0]; p->ph = d->pad[1]; p->pw = d->pad[2];
  M1_CHECK(d->nsrc >= 1 && d->nsrc <= M1_MAX_SRC, "conv: nsrc %d out of range", d->nsrc);
  M1_CHECK(d->nout >= 1 && d->nout <= M1_MAX_OUT, "conv: nout %d out of range", d->nout);
  p->nsrc = d->nsrc;
  p->nout = d->nout;
  p->n_total = 0;
  for (int s = 0; s < d->nsrc; ++s) p->src_c[s] = d->src_c[s];
  for (int j = 0; j < d->nout; ++j) {
    p->out_start[j] = p->n_total;
    p->out_c[j] = d->out_c[j];
    p->n_total += d->out_c[j];
  }
  static_assert(M1_MAX_OUT == M1_MAX_SRC, "stride arrays are shared between the two indexings");
  for (int j = 0; j < M1_MAX_OUT; ++j) {
    p->st[j] = d->w_stride_tap[j]; p->sr[j] = d->w_stride_red[j]; p->so[j] = d->w_stride_out[j];
  }
  p->w_by_src = d->w_by_src;
  p->out_start[d->nout] = p->n_total;
  p->accumulate = d->accumulate;
  p->out_vox = (int64_t)d->batch * p->Do * p->Ho * p->Wo;
  return 0;
}

}  // namespace

int m1_conv3d_simt(m1_ctx* ctx, const m1_conv_desc* d, const void* const* srcs,
                   const float* const* w, const float* const* bias, void* const* outs,
                   cudaStream_t st) {
  SimtParams p;
  if (fill_params(d, &p)) return 1;
  for (int s = 0; s < d->nsrc; ++s) p.src[s] = srcs[s];
  for (int j = 0; j < d->nout; ++j) {
    p.out[j] = outs[j];
    p.bias[j] = bias ? bias[j] : nullptr;
  }
  const int nw = d->w_by_src ? d->nout * d->nsrc : d->nout;
  for (int i = 0; i < nw; ++i) p.w[i] = w[i];
  const bool ib = d->act_dtype != M1_F32, ob = d->out_dtype != M1_F32;
  int k_total = 0;
  for (int s = 0; s < d->nsrc; ++s) k_total += d->src_c[s];
  const int taps = p.kd * p.kh * p.kw;
  const size_t thin_smem = (size_t)taps * k_total * TN * sizeof(float);
  if (pointwise_s1(d)) {
    HeadParams h;
    memset(&h, 0, sizeof(h));
    h.vox = p.out_vox;
    h.w = w[0];
    h.accumulate = d->accumulate & 1;
    if (!ob && p.n_total <= 8 && head_shape_ok(k_total, p.n_total)) {
      // features -> few fp32 channels
      h.x = srcs[0]; h.out = outs[0]; h.bias = bias ? bias[0] : nullptr;
      h.C = k_total; h.N = p.n_total; h.sr = d->w_stride_red[0]; h.so = d->w_stride_out[0];
      M1_DISPATCH_T(d->act_dtype, T, return launch_head<T>(ctx, 0, h, st));
    }
    if (!ib && k_total <= 8 && head_shape_ok(p.n_total, k_total) && !(bias && bias[0])) {
      // few fp32 gradient channels -> gradient of the features (roles of the weight strides swapped)
      h.dy = reinterpret_cast<const float*>(srcs[0]); h.out = outs[0];
      h.C = p.n_total; h.N = k_total; h.sr = d->w_stride_out[0]; h.so = d->w_stride_red[0];
      M1_DISPATCH_T(d->out_dtype, T, return launch_head<T>(ctx, 1, h, st));
    }
  }
  if ((k_total <= 16 || p.n_total <= 16) && thin_smem <= 48 * 1024) {
    dim3 tgrid((unsigned)cdiv64(p.out_vox, TV), (unsigned)((p.n_total + TN - 1) / TN));
    M1_DISPATCH_T2(d->act_dtype, d->out_dtype, T, TO, (conv_thin_kernel<T, TO><<<tgrid, 128, thin_smem, st>>>(p, k_total)));
    M1_LAUNCH_CHECK(ctx);
    return 0;
  }
  dim3 grid((unsigned)cdiv64(p.out_vox, BM), (unsigned)((p.n_total + BN - 1) / BN));
  M1_DISPATCH_T2(d->act_dtype, d->out_dtype, T, TO, (conv_simt_kernel<T, TO><<<grid, TH, 0, st>>>(p)));
  M1_LAUNCH_CHECK(ctx);
  return 0;
}

extern "C" int m1_conv3d_wgrad(m1_ctx* ctx, const m1_conv_desc* d, const void* const* srcs,
                               const void* const* douts, float* const* dw, float* const* dbias,
                               void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  WgradParams q;
  memset(&q, 0, sizeof(q));
  if (fill_params(d, &q.g)) return 1;
  for (int s = 0; s < d->nsrc; ++s) q.g.src[s] = srcs[s];
  const int taps = d->kernel[0] * d->kernel[1] * d->kernel[2];
  const int64_t out_vox = q.g.out_vox;
  // tensor-core engine: all produced tensors fused in ONE launch when the plan allows (the gathered
  // operand - the big one - is then streamed once), else one launch per produced tensor
  const bool fused_tc = d->engine != M1_ENGINE_SIMT && d->nout > 1 && m1_conv3d_wgrad_tc_supported(d, 0, d->nout);
  if (fused_tc) {
    if (m1_conv3d_wgrad_tc(ctx, d, 0, d->nout, srcs, douts, dw, st)) return 1;
  }
  for (int j = 0; j < d->nout; ++j) {
    q.dout = douts[j];
    q.Cn = d->out_c[j];
    q.dw = dw[j];
    q.st = d->w_stride_tap[j]; q.sr = d->w_stride_red[j]; q.so = d->w_stride_out[j];
    int r_base = 0;
    const bool use_tc = fused_tc || (d->engine != M1_ENGINE_SIMT && m1_conv3d_wgrad_tc_supported(d, j, 1));
    if (d->engine == M1_ENGINE_TCGEN05 && !use_tc) {
      m1_set_error("m1_conv3d_wgrad: launch not supported by the tcgen05 engine");
      return 1;
    }
    if (use_tc && !fused_tc) {
      if (m1_conv3d_wgrad_tc(ctx, d, j, 1, srcs, douts, dw, st)) return 1;
    }
    for (int s = 0; s < d->nsrc && !use_tc; ++s) {
      q.src_index = s;
      q.r_base = r_base;
      const int C = d->src_c[s];
      if (pointwise_s1(d) && d->out_dtype == M1_F32 && head_shape_ok(C, q.Cn)) {
        HeadParams h;
        memset(&h, 0, sizeof(h));
        h.vox = out_vox;
        h.x = srcs[s]; h.dy = reinterpret_cast<const float*>(douts[j]);
        h.dw = q.dw + (int64_t)r_base * q.sr;
        h.C = C; h.N = q.Cn; h.sr = q.sr; h.so = q.so;
        int rc = 0;
        M1_DISPATCH_T(d->act_dtype, T, rc = launch_head<T>(ctx, 2, h, st));
        if (rc) return 1;
        r_base += C;
        continue;
      }
      {
        // few-channel layer: exact-work kernel (micro-tiles of 4 x 8 outputs, <= 256 per block)
        const int Rp = (C + SW_RB - 1) / SW_RB * SW_RB, Np = (q.Cn + SW_NB - 1) / SW_NB * SW_NB;
        const int mt = (Rp / SW_RB) * (Np / SW_NB);
        if (mt <= 256 && (C < 48 || q.Cn < 48)) {
          int vch = 40960 / ((Rp + Np) * 4);
          vch = std::min(64, vch & ~7);
          if (vch >= 8) {
            int64_t want = std::max<int64_t>(1, (int64_t)ctx->num_sms * 8 / std::max(1, taps));
            int64_t vpb = std::max<int64_t>(cdiv64(cdiv64(out_vox, want), vch) * vch, (int64_t)vch * 4);
            q.vox_per_block = vpb;
            dim3 grid((unsigned)cdiv64(out_vox, vpb), (unsigned)taps);
            const size_t smem = (size_t)vch * (Rp + Np) * sizeof(float);
            M1_DISPATCH_T2(d->act_dtype, d->out_dtype, T, TO, (wgrad_small_kernel<T, TO><<<grid, 256, smem, st>>>(q, vch)));
            M1_LAUNCH_CHECK(ctx);
            r_base += C;
            continue;
          }
        }
      }
      const int tiles = ((C + BM - 1) / BM) * ((q.Cn + BN - 1) / BN);
      // split the voxel reduction so that the launch has ~8 blocks per SM
      int64_t want = std::max<int64_t>(1, (int64_t)ctx->num_sms * 8 / std::max(1, tiles * taps));
      int64_t vpb = cdiv64(out_vox, want);
      vpb = std::max<int64_t>(256, cdiv64(vpb, WV) * WV);
      q.vox_per_block = vpb;
      dim3 grid((unsigned)tiles, (unsigned)taps, (unsigned)cdiv64(out_vox, vpb));
      M1_DISPATCH_T2(d->act_dtype, d->out_dtype, T, TO, (wgrad_simt_kernel<T, TO><<<grid, TH, 0, st>>>(q)));
      M1_LAUNCH_CHECK(ctx);
      r_base += C;
    }
    if (dbias && dbias[j]) {
      const int C = q.Cn;
      const int64_t rpb = std::max<int64_t>(64, cdiv64(out_vox, (int64_t)ctx->num_sms * 4));
      const unsigned blocks = (unsigned)cdiv64(out_vox, rpb);
      M1_DISPATCH_T(d->out_dtype, T, (colsum_kernel<T><<<blocks, 256, C * sizeof(float), st>>>(
                                         reinterpret_cast<const T*>(douts[j]), out_vox, C, rpb, dbias[j])));
      M1_LAUNCH_CHECK(ctx);
    }
  }
  return 0;
}

extern "C" int m1_bias_grad(m1_ctx* ctx, const void* dout, int dtype, int64_t rows, int C, float* dbias,
                            void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t rpb = std::max<int64_t>(64, cdiv64(rows, (int64_t)ctx->num_sms * 4));
  const unsigned blocks = (unsigned)cdiv64(rows, rpb);
  M1_DISPATCH_T(dtype, T, (colsum_kernel<T><<<blocks, 256, C * sizeof(float), st>>>(reinterpret_cast<const T*>(dout),
                                                                                     rows, C, rpb, dbias)));
  M1_LAUNCH_CHECK(ctx);
  return 0;
}
