// augment.cu — K10: the reference's train-time augmentations on the device, batched, fp32.
//   tf2.5/scripts/model/augmentations.py:36-378 (`augment_tensors` and its helpers), applied per volume by
//   `train_gen.map(...)` on the host CPUs in the reference (train_model.py:181); here one launch per transform over
//   the whole batch, every sample with its own parameters (m1_aug_plan, drawn on the host).
// D plays the batch role of TensorFlow's 4-D image ops: every (sample, depth slice) is an independent H x W image.
// Third-party semantics restated (oracle/augment_oracle.py has the same statements on the CPU):
//   tf.image.resize bilinear / nearest : half-pixel centres, no antialiasing
//   tf.pad SYMMETRIC                   : reflection including the edge pixel
//   tfa.image.rotate BILINEAR          : output (x, y) reads input (cos x - sin y + x_off, sin x + cos y + y_off),
//                                        neighbours outside the image contribute 0
// All kernels: one thread per output element (b, d, h, w, c), NDHWC fp32; a sample whose transform is off is copied.
#include "common.cuh"

namespace {

constexpr int TB = 256;

struct Geo {
  int B, D, H, W, C;
  int64_t total;
};

__device__ __forceinline__ void decode(int64_t i, const Geo& g, int* b, int* d, int* h, int* w, int* c) {
  *c = (int)(i % g.C); i /= g.C;
  *w = (int)(i % g.W); i /= g.W;
  *h = (int)(i % g.H); i /= g.H;
  *d = (int)(i % g.D);
  *b = (int)(i / g.D);
}
__device__ __forceinline__ int64_t at(const Geo& g, int b, int d, int h, int w, int c) {
  return ((((int64_t)b * g.D + d) * g.H + h) * g.W + w) * g.C + c;
}
// index of a SYMMETRIC-padded axis back into [0, n)
__device__ __forceinline__ int reflect(int i, int n) { return i < 0 ? -1 - i : (i >= n ? 2 * n - 1 - i : i); }

// bilinear sample of the H x W slice (b, d, :, :, c) resized to (oh, ow), at output pixel (oy, ox)
__device__ __forceinline__ float resize_bilinear_at(const float* __restrict__ in, const Geo& g, int b, int d, int c,
                                                    int oh, int ow, int oy, int ox) {
  const float sy = ((float)oy + 0.5f) * ((float)g.H / (float)oh) - 0.5f;
  const float sx = ((float)ox + 0.5f) * ((float)g.W / (float)ow) - 0.5f;
  const float fy = floorf(sy), fx = floorf(sx);
  const int y0 = max((int)fy, 0), y1 = min((int)ceilf(sy), g.H - 1);
  const int x0 = max((int)fx, 0), x1 = min((int)ceilf(sx), g.W - 1);
  const float ly = sy - fy, lx = sx - fx;
  const float v00 = in[at(g, b, d, y0, x0, c)], v01 = in[at(g, b, d, y0, x1, c)];
  const float v10 = in[at(g, b, d, y1, x0, c)], v11 = in[at(g, b, d, y1, x1, c)];
  const float top = v00 + (v01 - v00) * lx, bot = v10 + (v11 - v10) * lx;
  return top + (bot - top) * ly;
}

__global__ void __launch_bounds__(TB) aug_zoom_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                      const m1_aug_plan* __restrict__ plans, Geo g) {
  const int64_t i = blockIdx.x * (int64_t)TB + threadIdx.x;
  if (i >= g.total) return;
  int b, d, h, w, c;
  decode(i, g, &b, &d, &h, &w, &c);
  const m1_aug_plan& p = plans[b];
  if (!p.zoom_on) { out[i] = in[i]; return; }
  const int S = p.zoom_scale;            // zoom_4D_tensor: resize to S x S, keep the bottom-right H x W window
  out[i] = resize_bilinear_at(in, g, b, d, c, S, S, h + (S - g.H), w + (S - g.W));
}

// which: 0 horizontal flip, 1 translation, 2 channel shift (translation of one of the first three channels)
__global__ void __launch_bounds__(TB) aug_remap_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                       const m1_aug_plan* __restrict__ plans, Geo g, int which) {
  const int64_t i = blockIdx.x * (int64_t)TB + threadIdx.x;
  if (i >= g.total) return;
  int b, d, h, w, c;
  decode(i, g, &b, &d, &h, &w, &c);
  const m1_aug_plan& p = plans[b];
  int sh = h, sw = w;
  if (which == 0) {
    if (p.flip_on) sw = g.W - 1 - w;
  } else if (which == 1) {
    // translate_4D_tensor: SYMMETRIC pad (top, bottom, left, right), then crop at (bottom, right)
    if (p.tr_on) { sh = reflect(h + p.tr_bottom - p.tr_top, g.H); sw = reflect(w + p.tr_right - p.tr_left, g.W); }
  } else {
    if (p.cs_on && c == p.cs_channel) {
      sh = reflect(h + p.cs_bottom - p.cs_top, g.H);
      sw = reflect(w + p.cs_right - p.cs_left, g.W);
    }
  }
  out[i] = in[at(g, b, d, sh, sw, c)];
}

__global__ void __launch_bounds__(TB) aug_rotate_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                        const m1_aug_plan* __restrict__ plans, Geo g) {
  const int64_t i = blockIdx.x * (int64_t)TB + threadIdx.x;
  if (i >= g.total) return;
  int b, d, h, w, c;
  decode(i, g, &b, &d, &h, &w, &c);
  const m1_aug_plan& p = plans[b];
  if (!p.rot_on) { out[i] = in[i]; return; }
  // rotate_4D_tensor: SYMMETRIC pad by rot_pad, rotate the padded image about its centre, central crop
  const int Hp = g.H + 2 * p.rot_pad, Wp = g.W + 2 * p.rot_pad;
  const float xp = (float)(w + p.rot_crop_w), yp = (float)(h + p.rot_crop_h);
  const float sx = p.rot_cos * xp - p.rot_sin * yp + p.rot_xoff;
  const float sy = p.rot_sin * xp + p.rot_cos * yp + p.rot_yoff;
  const float x0 = floorf(sx), y0 = floorf(sy);
  auto read = [&](float yy, float xx) -> float {
    if (yy < 0.f || yy >= (float)Hp || xx < 0.f || xx >= (float)Wp) return 0.f;
    return in[at(g, b, d, reflect((int)yy - p.rot_pad, g.H), reflect((int)xx - p.rot_pad, g.W), c)];
  };
  const float wx1 = sx - x0, wy1 = sy - y0, wx0 = x0 + 1.f - sx, wy0 = y0 + 1.f - sy;
  const float top = wx0 * read(y0, x0) + wx1 * read(y0, x0 + 1.f);
  const float bot = wx0 * read(y0 + 1.f, x0) + wx1 * read(y0 + 1.f, x0 + 1.f);
  out[i] = wy0 * top + wy1 * bot;
}

// sim_poor_scan_3D_tensor per channel: bilinear down to (L, L), L = int(0.75 H), nearest back up to (H, H)
__global__ void __launch_bounds__(TB) aug_poor_scan_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                           const m1_aug_plan* __restrict__ plans, Geo g) {
  const int64_t i = blockIdx.x * (int64_t)TB + threadIdx.x;
  if (i >= g.total) return;
  int b, d, h, w, c;
  decode(i, g, &b, &d, &h, &w, &c);
  const m1_aug_plan& p = plans[b];
  if (c >= 3 || !p.poor_on[c]) { out[i] = in[i]; return; }
  const int L = (int)((float)g.H * 0.75f);
  const int ly = min((int)floorf(((float)h + 0.5f) * ((float)L / (float)g.H)), L - 1);
  const int lx = min((int)floorf(((float)w + 0.5f) * ((float)L / (float)g.W)), L - 1);
  out[i] = resize_bilinear_at(in, g, b, d, c, L, L, ly, lx);
}

// gaussian_noise_4D_tensor: x[..., :3] += stddev * eps (eps ~ N(0,1), (B, D, H, W, 3))
__global__ void __launch_bounds__(TB) aug_noise_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                       const float* __restrict__ eps,
                                                       const m1_aug_plan* __restrict__ plans, Geo g) {
  const int64_t i = blockIdx.x * (int64_t)TB + threadIdx.x;
  if (i >= g.total) return;
  int b, d, h, w, c;
  decode(i, g, &b, &d, &h, &w, &c);
  const m1_aug_plan& p = plans[b];
  float v = in[i];
  if (p.noise_on && c < 3) v += p.noise_std * eps[((((int64_t)b * g.D + d) * g.H + h) * g.W + w) * 3 + c];
  out[i] = v;
}

// ---- gamma_shift_3D_tensor: per (sample, channel < 3) statistics over the whole volume, three passes -------------
// stats1[b][c] = {min, max, mean, std} of x; stats2[b][c] = {mean, std} of the transformed x_.
// One block per (sample, channel): the reduction order is fixed (deterministic), the data is 2 M elements.
template <int K>
__device__ __forceinline__ void block_reduce_d(double (&v)[K], double* sm) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < K; ++k)
    for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
  if (lane == 0)
#pragma unroll
    for (int k = 0; k < K; ++k) sm[k * (TB / 32) + warp] = v[k];
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
      double t = lane < TB / 32 ? sm[k * (TB / 32) + lane] : 0.0;
      for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
      v[k] = t;
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(TB) aug_gamma_stats1_kernel(const float* __restrict__ in,
                                                              const m1_aug_plan* __restrict__ plans, Geo g,
                                                              float* __restrict__ stats1) {
  __shared__ double sm[2 * (TB / 32)];
  __shared__ float smm[2 * (TB / 32)];
  const int b = blockIdx.x / 3, c = blockIdx.x % 3;
  if (c >= g.C || !plans[b].gamma_on[c]) return;
  const int64_t vox = (int64_t)g.D * g.H * g.W;
  const float* x = in + (int64_t)b * vox * g.C + c;
  double s[2] = {0.0, 0.0};
  float lo = INFINITY, hi = -INFINITY;
  for (int64_t v = threadIdx.x; v < vox; v += TB) {
    const float t = x[v * g.C];
    s[0] += t; s[1] += (double)t * t;
    lo = fminf(lo, t); hi = fmaxf(hi, t);
  }
  block_reduce_d<2>(s, sm);
  for (int o = 16; o > 0; o >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0) { smm[threadIdx.x >> 5] = lo; smm[TB / 32 + (threadIdx.x >> 5)] = hi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < TB / 32; ++k) { lo = fminf(lo, smm[k]); hi = fmaxf(hi, smm[TB / 32 + k]); }
    const double mean = s[0] / (double)vox, var = fmax(s[1] / (double)vox - mean * mean, 0.0);
    float* o = stats1 + (b * 3 + c) * 4;
    o[0] = lo; o[1] = hi; o[2] = (float)mean; o[3] = (float)sqrt(var);
  }
}

__device__ __forceinline__ float gamma_core(float t, float lo, float hi, float gamma) {
  return powf((t - lo) / (hi - lo + 1e-8f), gamma) * (hi - lo) + lo;
}

__global__ void __launch_bounds__(TB) aug_gamma_stats2_kernel(const float* __restrict__ in,
                                                              const m1_aug_plan* __restrict__ plans, Geo g,
                                                              const float* __restrict__ stats1,
                                                              float* __restrict__ stats2) {
  __shared__ double sm[2 * (TB / 32)];
  const int b = blockIdx.x / 3, c = blockIdx.x % 3;
  if (c >= g.C || !plans[b].gamma_on[c]) return;
  const int64_t vox = (int64_t)g.D * g.H * g.W;
  const float* x = in + (int64_t)b * vox * g.C + c;
  const float lo = stats1[(b * 3 + c) * 4], hi = stats1[(b * 3 + c) * 4 + 1], gm = plans[b].gamma;
  double s[2] = {0.0, 0.0};
  for (int64_t v = threadIdx.x; v < vox; v += TB) {
    const float t = gamma_core(x[v * g.C], lo, hi, gm);
    s[0] += t; s[1] += (double)t * t;
  }
  block_reduce_d<2>(s, sm);
  if (threadIdx.x == 0) {
    const double mean = s[0] / (double)vox, var = fmax(s[1] / (double)vox - mean * mean, 0.0);
    stats2[(b * 3 + c) * 2] = (float)mean;
    stats2[(b * 3 + c) * 2 + 1] = (float)sqrt(var);
  }
}

__global__ void __launch_bounds__(TB) aug_gamma_apply_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                             const m1_aug_plan* __restrict__ plans, Geo g,
                                                             const float* __restrict__ stats1,
                                                             const float* __restrict__ stats2) {
  const int64_t i = blockIdx.x * (int64_t)TB + threadIdx.x;
  if (i >= g.total) return;
  const int c = (int)(i % g.C);
  const int b = (int)(i / ((int64_t)g.D * g.H * g.W * g.C));
  const m1_aug_plan& p = plans[b];
  float v = in[i];
  if (c < 3 && p.gamma_on[c]) {
    const float* s1 = stats1 + (b * 3 + c) * 4;
    const float* s2 = stats2 + (b * 3 + c) * 2;
    float t = gamma_core(v, s1[0], s1[1], p.gamma) - s2[0];
    t = t / (s2[1] + 1e-8f) * s1[3];            // retain the original intensity distribution shape
    v = t + s1[2];
  }
  out[i] = v;
}

}  // namespace

extern "C" int m1_augment(m1_ctx* ctx, int op, const float* in, float* out, const float* eps,
                          const m1_aug_plan* plans, int batch, int D, int H, int W, int C, void* stream) {
  M1_CHECK(ctx && in && out && plans && in != out, "m1_augment: NULL or aliased argument");
  M1_CHECK(batch > 0 && D > 0 && H > 0 && W > 0 && C > 0, "m1_augment: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  Geo g{batch, D, H, W, C, (int64_t)batch * D * H * W * C};
  const unsigned blocks = (unsigned)((g.total + TB - 1) / TB);
  switch (op) {
    case M1_AUG_ZOOM: aug_zoom_kernel<<<blocks, TB, 0, st>>>(in, out, plans, g); break;
    case M1_AUG_FLIP: aug_remap_kernel<<<blocks, TB, 0, st>>>(in, out, plans, g, 0); break;
    case M1_AUG_TRANSLATE: aug_remap_kernel<<<blocks, TB, 0, st>>>(in, out, plans, g, 1); break;
    case M1_AUG_CHANNEL_SHIFT: aug_remap_kernel<<<blocks, TB, 0, st>>>(in, out, plans, g, 2); break;
    case M1_AUG_ROTATE: aug_rotate_kernel<<<blocks, TB, 0, st>>>(in, out, plans, g); break;
    case M1_AUG_POOR_SCAN:
      M1_CHECK(H == W, "m1_augment: sim_poor_scan resizes to (H, H) like the reference and needs square slices");
      aug_poor_scan_kernel<<<blocks, TB, 0, st>>>(in, out, plans, g);
      break;
    case M1_AUG_NOISE:
      M1_CHECK(eps != nullptr, "m1_augment: the noise transform needs the N(0,1) tensor (B, D, H, W, 3)");
      aug_noise_kernel<<<blocks, TB, 0, st>>>(in, out, eps, plans, g);
      break;
    case M1_AUG_GAMMA: {
      M1_CHECK((size_t)batch * 3 * 6 * sizeof(float) <= ctx->scratch_bytes, "m1_augment: scratch too small");
      float* s1 = ctx->scratch;
      float* s2 = s1 + (size_t)batch * 3 * 4;
      aug_gamma_stats1_kernel<<<batch * 3, TB, 0, st>>>(in, plans, g, s1);
      M1_LAUNCH_CHECK(ctx);
      aug_gamma_stats2_kernel<<<batch * 3, TB, 0, st>>>(in, plans, g, s1, s2);
      M1_LAUNCH_CHECK(ctx);
      aug_gamma_apply_kernel<<<blocks, TB, 0, st>>>(in, out, plans, g, s1, s2);
      break;
    }
    default: M1_CHECK(false, "m1_augment: unknown transform %d", op);
  }
  M1_LAUNCH_CHECK(ctx);
  return 0;
}
