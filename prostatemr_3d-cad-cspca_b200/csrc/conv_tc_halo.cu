// conv_tc_halo.cu — K1/K2, stride-1 gathers with in-plane filter taps (1x3x3 / 3x3x3 Conv3D, their data
// gradients): implicit GEMM on tcgen05 where ONE halo'd activation tile in shared memory serves all
// in-plane taps.
//
// conv_tc.cu loads one 128-voxel TMA box per (tap, channel chunk): every activation byte crosses L2->SM
// kh*kw times, and TMA throughput is per LINE (<= 128 B inner run, ~2 cycles each), so the few-channel /
// full-resolution layers of M1 (R:network_blocks.py:37-46 at 160x160 / 80x80) are bound by the number of
// boxes they issue, not by the tensor cores. Here the M tile is 128 consecutive rows of a LINEARISED halo tile:
//   - the A box of a (plane, channel chunk) is the (L = G*bh + kh-1) x (P = bw + kw-1) voxel halo
//     rectangle around G stacked sub-tiles of bh x bw output voxels, rows in (line, column) order,
//     P rows per line, canonical K-major swizzled layout as written by TMA (OOB zero fill == SAME pad);
//   - output row q of sub-tile g is voxel (line g*bh + q / P, column q % P); its input for tap (b, c)
//     is row q + g*bh*P + sb*P + sc of the same tile: a pure ROW SHIFT of the UMMA descriptor start
//     address. tcgen05 applies the shared-memory swizzle on absolute address bits, so a descriptor may
//     start at any row of a TMA-written tile (tools/probe_umma_shift.cu verifies this on B200 for
//     SW128/64/32, K-major and MN-major);
//   - rows with q % P >= bw or q / P >= bh are halo positions: their accumulator rows are garbage and are
//     never stored (MMA rows are independent).
// Activation traffic per (plane, chunk) drops from kh*kw*128 rows to L*P rows (9 x -> ~1.5 x), the weight
// tile of a tap is shared by the G sub-tiles (G accumulators in TMEM).
#include "tc_epilogue.cuh"
#include <algorithm>
#include <cstdlib>

namespace {
using namespace tc;

constexpr int kThreads = 256;   // warp 0: TMA producer, warp 1: MMA issuer, all 8 warps: epilogue

struct HaloParams {
  CUtensorMap tmA[M1_MAX_SRC];
  CUtensorMap tmB;
  int nsrc;
  int src_chunks[M1_MAX_SRC];
  int src_koff[M1_MAX_SRC];
  int kd, khw;                 // planes, in-plane taps
  int8_t plane_off[3];         // gathered-grid d offset of plane a
  uint16_t row_shift[9];       // in-plane tap t -> row shift sb*P + sc
  int lo_h, lo_w;              // halo origin relative to the first output voxel of the tile
  int G, bh, bw, P, L;
  int th, tw;                  // tiles per plane
  int D, H, W;                 // grid (gathered == produced for stride 1)
  int n_tile, ck;
  int stages;
  uint32_t a_alloc, b_tile_bytes, stage_bytes, tx_bytes;
  uint32_t tmem_cols;
  uint32_t idesc, desc_hi;
  EpiOut epi;
  int dbg;                     // timing experiments: 1 = no A loads, 2 = no B loads, 4 = no MMAs (results invalid)
};

__global__ void __launch_bounds__(kThreads) conv_halo_kernel(const __grid_constant__ HaloParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_full = smem_base;
  const uint32_t bar_empty = smem_base + 8u * 16u;
  const uint32_t bar_accum = smem_base + 8u * 32u;
  const uint32_t tmem_slot = smem_base + 8u * 33u;
  const uint32_t tiles = smem_base + 1024u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5;

  int t = blockIdx.x;
  const int tw_i = t % p.tw; t /= p.tw;
  const int th_i = t % p.th; t /= p.th;
  const int d = t % p.D; t /= p.D;
  const int n_img = t;
  const int h0 = th_i * p.G * p.bh, w0 = tw_i * p.bw;
  const int n0 = blockIdx.y * p.n_tile;
  // sub-tiles that start inside the volume (the last tile of a column may own fewer)
  const int g_live = min(p.G, (p.H - h0 + p.bh - 1) / p.bh);

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                 "r"(p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 32) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(bar_full + 8u * s, 1);
      mbar_init(bar_empty + 8u * s, 1);
    }
    mbar_init(bar_accum, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + 8u * 33u);

  int total_chunks = 0;
  for (int s = 0; s < p.nsrc; ++s) total_chunks += p.src_chunks[s];
  // planes whose gathered d lies inside the volume (the others contribute zeros: skipped)
  int plane_mask = 0, nplanes = 0;
  for (int a = 0; a < p.kd; ++a) {
    const int di = d + p.plane_off[a];
    if (di >= 0 && di < p.D) { plane_mask |= 1 << a; ++nplanes; }
  }
  const int iters = nplanes * total_chunks;

  if (warp == 0) {
    if (elect_one()) {
      // ===== TMA producer: one stage = halo box of one (plane, channel chunk) + the weight tiles of its taps
      uint32_t stage = 0, phase = 0;
      for (int a = 0; a < p.kd; ++a) {
        if (!((plane_mask >> a) & 1)) continue;
        for (int src = 0; src < p.nsrc; ++src)
          for (int chunk = 0; chunk < p.src_chunks[src]; ++chunk) {
            mbar_wait(bar_empty + 8u * stage, phase ^ 1u);
            const uint32_t full = bar_full + 8u * stage;
            const uint32_t a_bytes = (uint32_t)(p.L * p.P) * p.ck * 2u;
            mbar_expect_tx(full, ((p.dbg & 1) ? 0u : a_bytes) + ((p.dbg & 2) ? 0u : p.tx_bytes - a_bytes));
            const uint32_t sbase = tiles + stage * p.stage_bytes;
            if (!(p.dbg & 1))
              tma_load_5d(sbase, &p.tmA[src], full, chunk * p.ck, w0 + p.lo_w, h0 + p.lo_h, d + p.plane_off[a], n_img);
            if (!(p.dbg & 2))
              tma_load_3d(sbase + p.a_alloc, &p.tmB, full, p.src_koff[src] + chunk * p.ck, n0, a * p.khw);
            if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1u; }
          }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (elect_one()) {
      const uint64_t hi = (uint64_t)p.desc_hi << 32;
      // ===== MMA issuer. The loop body is kept to a handful of uniform-datapath adds per MMA: a tensor-core
      // instruction of this shape costs only ~40-80 cycles, address arithmetic in the issuing thread is exposed.
      uint32_t stage = 0, phase = 0;
      const uint32_t row16 = (2u * p.ck) >> 4;                   // row pitch in 16-byte units
      const uint32_t b_tile16 = p.b_tile_bytes >> 4;
      uint32_t g_off[4], g_tmem[4];
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        g_off[g] = (uint32_t)(g * p.bh * p.P) * row16;
        g_tmem[g] = tmem_base + (uint32_t)(g * p.n_tile);
      }
      const int k16s = p.ck / 16;
      for (int it = 0; it < iters; ++it) {
        mbar_wait(bar_full + 8u * stage, phase);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a0 = (tiles + stage * p.stage_bytes) >> 4;      // < 2^14: shared memory is < 256 KB
        const uint32_t b0 = a0 + (p.a_alloc >> 4);
        if (!(p.dbg & 4)) {
          for (int tp = 0; tp < p.khw; ++tp) {
            const uint32_t at = a0 + (uint32_t)p.row_shift[tp] * row16;
            const uint32_t bt = b0 + (uint32_t)tp * b_tile16;
            const uint32_t first = (it | tp) ? 1u : 0u;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              if (g < g_live) {
                const uint32_t ag = at + g_off[g];
                umma_f16(g_tmem[g], hi | ag, hi | bt, p.idesc, first);
                if (k16s > 1) umma_f16(g_tmem[g], hi | (ag + 2u), hi | (bt + 2u), p.idesc, 1u);
                if (k16s > 2) {
                  umma_f16(g_tmem[g], hi | (ag + 4u), hi | (bt + 4u), p.idesc, 1u);
                  umma_f16(g_tmem[g], hi | (ag + 6u), hi | (bt + 6u), p.idesc, 1u);
                }
              }
            }
          }
        }
        umma_commit(bar_empty + 8u * stage);
        if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1u; }
      }
      umma_commit(bar_accum);
    }
    __syncwarp();
  }

  // ===== epilogue: 8 warps; warps w and w+4 share TMEM lanes 32*(w%4).. and split the columns
  mbar_wait(bar_accum, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  {
    const int q = (warp & 3) * 32 + (threadIdx.x & 31);       // accumulator row == TMEM lane
    const int lh = q / p.P, lw = q % p.P;
    const int w = w0 + lw;
    const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    int cb, ce;
    epi_cols(warp, p.n_tile, &cb, &ce);
    for (int g = 0; g < g_live; ++g) {
      const int h = h0 + g * p.bh + lh;
      const bool valid = lh < p.bh && lw < p.bw && h < p.H && w < p.W;
      const int64_t vox = (((int64_t)n_img * p.D + d) * p.H + h) * p.W + w;
      epilogue_row(p.epi, lane_addr + (uint32_t)(g * p.n_tile), n0, cb, ce, valid, vox, iters == 0);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols)
                 : "memory");
  }
}

struct HaloPlan {
  int ck, n_real, n_total, k_total, n_tile, n_tiles;
  int G, bh, bw, P, L, th, tw;
  int stages, ctas_per_sm;
  uint32_t a_alloc, b_tile_bytes, stage_bytes, smem_bytes, tmem_cols;
  double bytes_per_voxel;     // L2->SM traffic model of this plan
};

bool make_halo_plan(const m1_conv_desc* d, HaloPlan* pl) {
  static const int enabled = getenv("M1_HALO") ? atoi(getenv("M1_HALO")) : 1;
  if (!enabled) return false;
  if (!m1_is16(d->act_dtype) || !m1_is16(d->out_dtype)) return false;
  // one operand format per MMA: kind::f16 with different A / B formats is an illegal instruction on sm_100
  if (d->w_dtype != 0 && d->w_dtype != d->act_dtype) return false;
  if (d->nsrc < 1 || d->nsrc > M1_MAX_SRC || d->nout < 1 || d->nout > M1_MAX_OUT) return false;
  for (int i = 0; i < 3; ++i) {
    if (d->stride[i] != 1) return false;
    if (d->kernel[i] != 1 && d->kernel[i] != 3) return false;
    if (d->pad[i] != (d->kernel[i] - 1) / 2) return false;
    if (d->in_dhw[i] != d->out_dhw[i]) return false;
  }
  const int kh = d->kernel[1], kw = d->kernel[2];
  if (kh * kw == 1) return false;            // nothing to share
  int ck = 64, k_total = 0;
  for (int s = 0; s < d->nsrc; ++s) {
    const int c = d->src_c[s];
    if (c % 16) return false;
    while (c % ck) ck >>= 1;
    k_total += c;
  }
  int n_total = 0;
  for (int j = 0; j < d->nout; ++j) {
    if (d->out_c[j] % 8) return false;
    n_total += d->out_c[j];
  }
  const int n_real = n_total;
  n_total = (n_total + 15) & ~15;
  int n_tile = 0;
  for (int c = 256; c >= 16; c -= 16)
    if (n_total % c == 0) { n_tile = c; break; }
  if (!n_tile) return false;
  const int H = d->out_dhw[1], W = d->out_dhw[2];
  const int khw = kh * kw;
  static const int g_force = getenv("M1_HALO_G") ? atoi(getenv("M1_HALO_G")) : 0;
  static const int ck_force = getenv("M1_HALO_CK") ? atoi(getenv("M1_HALO_CK")) : 0;
  bool found = false;
  double best = 1e30;
  pl->ctas_per_sm = 1;
  for (int cc = ck; cc >= 16; cc >>= 1) {
    if (ck_force && cc != ck_force && ck_force <= ck) continue;
    const uint32_t row_bytes = 2u * cc;
    const uint32_t b_tile = (uint32_t)n_tile * row_bytes;
    for (int G = 4; G >= 1; --G) {
      if (g_force && G != g_force) continue;
      if (G * n_tile > 512) continue;
      uint32_t cols = 32;
      while ((int)cols < G * n_tile) cols <<= 1;
      for (int bw = std::min(W, 128); bw >= 4; --bw) {
        const int P = bw + kw - 1;
        const int bh = std::min(H, (128 - bw) / P + 1);
        if (bh < 1) continue;
        const int L = G * bh + kh - 1;
        if (P > 256 || L > 256) continue;
        // rows touched by the last shifted MMA of the last sub-tile
        const int rows = std::max(L * P, (G - 1) * bh * P + (kh - 1) * P + (kw - 1) + 128);
        const uint32_t a_alloc = ((uint32_t)rows * row_bytes + 1023u) & ~1023u;
        const uint32_t stage = a_alloc + (((uint32_t)khw * b_tile + 1023u) & ~1023u);
        if (2u * stage + 2048u > 227u * 1024u) continue;
        int ctas = std::min<int>(512 / (int)cols, (int)((227u * 1024u) / (2u * stage + 2048u)));
        ctas = std::max(1, std::min(ctas, 4));
        const int th = (H + G * bh - 1) / (G * bh), tw = (W + bw - 1) / bw;
        const double vox_eff = (double)H * W / ((double)th * tw * G * bh * bw);   // useful / owned
        const double mma_eff = (double)bh * bw / 128.0 * vox_eff;
        // L2->SM bytes per useful output voxel and channel chunk of 16 (A rows + weights)
        const double bytes = ((double)L * P + (double)khw * n_tile) / ((double)G * bh * bw * vox_eff);
        // The kernel is bound by the MMA issue rate (~45 cycles per M=128 x K=16 instruction at small N), not by
        // traffic: first maximise the useful fraction of the MMA rows and keep >= 2 CTAs per SM (the epilogue of
        // one overlaps the main loop of the other); among equals prefer wide channel chunks (fewer, longer TMA
        // lines and barrier round trips), then low traffic (large G).
        const double score = 1.0 / (mma_eff * (ctas >= 2 ? 1.0 : 0.75)) + 0.02 * (64.0 / cc) + 0.002 * bytes;
        if (score < best) {
          best = score; found = true;
          pl->ck = cc; pl->G = G; pl->bh = bh; pl->bw = bw; pl->P = P; pl->L = L; pl->th = th; pl->tw = tw;
          pl->a_alloc = a_alloc; pl->b_tile_bytes = b_tile; pl->stage_bytes = stage; pl->tmem_cols = cols;
          pl->ctas_per_sm = ctas;
          pl->bytes_per_voxel = bytes;
        }
      }
    }
  }
  if (!found) return false;
  pl->n_real = n_real; pl->n_total = n_total; pl->k_total = k_total; pl->n_tile = n_tile;
  pl->n_tiles = n_total / n_tile;
  const uint32_t budget = (227u * 1024u) / pl->ctas_per_sm - 2048u;
  int stages = (int)(budget / pl->stage_bytes);
  const int iters = d->kernel[0] * (k_total / pl->ck);
  stages = std::max(2, std::min(stages, std::min(6, std::max(2, iters))));
  pl->stages = stages;
  pl->smem_bytes = 2048u + (uint32_t)stages * pl->stage_bytes;
  if (pl->smem_bytes > 227u * 1024u) return false;
  return true;
}

}  // namespace

// 1 if the halo engine can run the launch; *preferred = 1 if the built-in heuristic would pick it over the
// per-tap engine (few produced channels: the per-tap engine is bound by its many small TMA boxes there,
// while wide-N layers on coarse grids are already tensor-bound and would only lose MMA rows to the halo
// columns). The host may override the choice per layer shape after timing both (m1_conv_desc.tune[0]).
int m1_conv3d_halo_supported(const m1_conv_desc* d, int* preferred) {
  HaloPlan pl;
  if (!make_halo_plan(d, &pl)) return 0;
  if (preferred) {
    const double mma_eff = (double)pl.bh * pl.bw / 128.0;
    *preferred = (pl.n_tile <= 96 && mma_eff >= 0.85) ? 1 : 0;
  }
  return 1;
}

// out: ck, n_tile, n_tiles, G, bh, bw, P, L, stages, smem_bytes, tmem_cols, ctas_per_sm, stage_bytes, a_alloc
int m1_conv3d_halo_plan_info(const m1_conv_desc* d, int32_t* out) {
  HaloPlan pl;
  if (!make_halo_plan(d, &pl)) return 0;
  const int32_t v[14] = {pl.ck, pl.n_tile, pl.n_tiles, pl.G, pl.bh, pl.bw, pl.P, pl.L, pl.stages,
                         (int32_t)pl.smem_bytes, (int32_t)pl.tmem_cols, pl.ctas_per_sm, (int32_t)pl.stage_bytes,
                         (int32_t)pl.a_alloc};
  for (int i = 0; i < 14; ++i) out[i] = v[i];
  return 14;
}

int m1_conv3d_halo(m1_ctx* ctx, const m1_conv_desc* d, const void* const* srcs, const void* w_packed,
                   const float* const* bias, void* const* outs, cudaStream_t st) {
  HaloPlan pl;
  M1_CHECK(make_halo_plan(d, &pl), "m1_conv3d: launch not supported by the halo tcgen05 engine");
  M1_CHECK(w_packed != nullptr, "m1_conv3d: tcgen05 engine needs the bf16 weight pack");
  M1_CHECK(ctx->encode_tiled != nullptr, "cuTensorMapEncodeTiled entry point unavailable");
  EncodeTiledFn encode = reinterpret_cast<EncodeTiledFn>(ctx->encode_tiled);
  static_assert(sizeof(HaloParams) < 4000, "kernel parameter block too large");
  HaloParams p;
  memset(&p, 0, sizeof(p));
  const int D = d->out_dhw[0], H = d->out_dhw[1], W = d->out_dhw[2];
  const int kd = d->kernel[0], kh = d->kernel[1], kw = d->kernel[2];
  const bool fwd = d->mode == M1_CONV_FWD;
  int koff = 0;
  for (int s = 0; s < d->nsrc; ++s) {
    M1_CHECK(((uintptr_t)srcs[s] & 15) == 0, "m1_conv3d: gathered tensor %d not 16-byte aligned", s);
    int r = encode_ndhwc(encode, &p.tmA[s], srcs[s], d->act_dtype, d->src_c[s], W, H, D, d->batch, pl.ck, pl.P, pl.L, 1);
    M1_CHECK(r == 0, "cuTensorMapEncodeTiled(halo A %d) failed: %d", s, r);
    p.src_chunks[s] = d->src_c[s] / pl.ck;
    p.src_koff[s] = koff;
    koff += d->src_c[s];
  }
  const int taps = kd * kh * kw;
  {
    cuuint64_t dims[3] = {(cuuint64_t)pl.k_total, (cuuint64_t)pl.n_total, (cuuint64_t)taps};
    cuuint64_t strides[2] = {(cuuint64_t)pl.k_total * 2, (cuuint64_t)pl.k_total * 2 * pl.n_total};
    cuuint32_t box[3] = {(cuuint32_t)pl.ck, (cuuint32_t)pl.n_tile, (cuuint32_t)(kh * kw)};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = encode(&p.tmB, tm_dtype(m1_conv_w_dtype(d)), 3, const_cast<void*>(w_packed), dims, strides,
                        box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(pl.ck), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    M1_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(halo B) failed: %d", (int)r);
  }
  p.nsrc = d->nsrc;
  p.kd = kd; p.khw = kh * kw;
  // tap k reads gathered voxel o + (k - pad) [FWD] or o + (pad - k) [TRANSPOSED, stride 1]
  for (int a = 0; a < kd; ++a) p.plane_off[a] = (int8_t)(fwd ? a - d->pad[0] : d->pad[0] - a);
  p.lo_h = fwd ? -d->pad[1] : d->pad[1] - (kh - 1);
  p.lo_w = fwd ? -d->pad[2] : d->pad[2] - (kw - 1);
  for (int b = 0; b < kh; ++b)
    for (int c = 0; c < kw; ++c) {
      const int sb = (fwd ? b - d->pad[1] : d->pad[1] - b) - p.lo_h;
      const int sc = (fwd ? c - d->pad[2] : d->pad[2] - c) - p.lo_w;
      p.row_shift[b * kw + c] = (uint16_t)(sb * pl.P + sc);
    }
  p.G = pl.G; p.bh = pl.bh; p.bw = pl.bw; p.P = pl.P; p.L = pl.L;
  p.th = pl.th; p.tw = pl.tw;
  p.D = D; p.H = H; p.W = W;
  p.n_tile = pl.n_tile; p.ck = pl.ck;
  p.stages = pl.stages;
  p.a_alloc = pl.a_alloc; p.b_tile_bytes = pl.b_tile_bytes; p.stage_bytes = pl.stage_bytes;
  p.tx_bytes = (uint32_t)(pl.L * pl.P) * pl.ck * 2u + (uint32_t)(kh * kw) * pl.b_tile_bytes;
  p.tmem_cols = pl.tmem_cols;
  p.idesc = (1u << 4) | (idesc_fmt(d->act_dtype) << 7) | (idesc_fmt(m1_conv_w_dtype(d)) << 10) |
            ((uint32_t)(pl.n_tile >> 3) << 17) | ((128u >> 4) << 24);
  p.desc_hi = ((8u * pl.ck * 2u) >> 4) | (1u << 14) | (layout_for(pl.ck) << 29);
  M1_CHECK(epi_fill(&p.epi, d, bias, outs), "m1_conv3d: too many produced channels for the tcgen05 epilogue table");
  for (int j = 0; j < d->nout; ++j)
    M1_CHECK(((uintptr_t)outs[j] & 15) == 0, "m1_conv3d: produced tensor %d not 16-byte aligned", j);
  static const int g_dbg = getenv("M1_HALO_DBG") ? atoi(getenv("M1_HALO_DBG")) : 0;
  p.dbg = g_dbg;
  if (getenv("M1_HALO_VERBOSE"))
    fprintf(stderr, "halo plan: ck %d G %d bh %d bw %d P %d L %d stages %d stage_bytes %u ctas/SM %d tmem %u grid %d x %d\n",
            pl.ck, pl.G, pl.bh, pl.bw, pl.P, pl.L, pl.stages, pl.stage_bytes, pl.ctas_per_sm, pl.tmem_cols,
            d->batch * D * pl.th * pl.tw, pl.n_tiles);
  const uint32_t smem_bytes = pl.smem_bytes;
  static int smem_set = 0;
  if (!smem_set) {
    M1_CUDA(cudaFuncSetAttribute(conv_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    smem_set = 1;
  }
  dim3 grid((unsigned)(d->batch * D * pl.th * pl.tw), (unsigned)pl.n_tiles);
  conv_halo_kernel<<<grid, kThreads, smem_bytes, st>>>(p);
  M1_LAUNCH_CHECK(ctx);
  return 0;
}
