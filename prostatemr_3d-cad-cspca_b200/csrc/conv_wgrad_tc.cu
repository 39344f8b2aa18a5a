// conv_wgrad_tc.cu — K3: weight gradient of a stride-1 3-D convolution on the tcgen05 tensor cores.
//
// Replaces Conv3DBackpropFilterV2 (autodiff of tf.keras.layers.Conv3D, R:network_blocks.py:37-46).
//
//   dW[tap][r][n] = sum over output voxels o of  X[o + tap - pad][r] * dY[o][n]
//
// GEMM view per tap:   D[M = 128 reduced channels r, N = produced channels] += A[M, K] * B[K, N]
// with K = output voxels.  Both operands are "MN-major" for the tensor core: their contraction index
// (the voxel) is the slow one in memory, channels are contiguous - exactly how NDHWC tensors lie in
// HBM, so again TMA boxes feed the MMA with no data movement by threads:
//   A block  = 5-D TMA box (ck channels, bw, bh, bd, 1) of one gathered tensor, corner shifted by the
//              tap (zero fill == SAME padding); 128/ck such blocks (possibly from DIFFERENT tensors of
//              the virtual concatenation) form the M = 128 rows, LBO = block bytes
//   B block  = 5-D TMA box (cb channels, bw, bh, bd, 1) of dY, n_tile/cb blocks form N
//   K step   = 16 voxels = 16 rows of every block (2 swizzle groups of 8 rows, SBO = 8 rows)
// A CTA owns one (tap group, M tile, N tile) and a slab of voxel bricks (split-K); the taps of a
// group differ only in kw and share the B blocks (one accumulator per tap in TMEM); partial dW tiles
// are added to the fp32 gradient with atomics (shared weights accumulate over passes anyway).
//
// SHIFT mode (m1_conv_desc.tune[1] == 2; stride 1, kw == 3): the three kw taps of a group also share ONE
// activation box. A brick is bh full-width lines loaded with a row pitch P >= W + 2 for BOTH operands
// (bh * P % 16 == 0): the X box starts at w = -1, the dY box at w = 0, columns beyond the volume are
// zero-filled by TMA. Row q of dY then pairs with row q + c of X for tap c - a row shift of the MN-major
// UMMA descriptor (tcgen05 swizzles on absolute address bits, tools/probe_umma_shift.cu). dY rows whose
// column is >= W are zero, so the halo rows of X (and the up to two rows read past a block, kept finite by
// zeroed pad rows) contribute nothing. TMA lines of X per brick: one third of the per-tap scheme.
#include "tc_common.cuh"
#include <algorithm>
#include <cmath>
#include <cstdlib>

namespace {
using namespace tc;

constexpr int kThreads = 128;
constexpr int kMaxBlocks = 64;   // 16-channel blocks over the concatenation handled per launch (<= 1024 ch)

struct WgParams {
  CUtensorMap tmA[M1_MAX_SRC];
  CUtensorMap tmB[M1_MAX_OUT];
  int nsrc;
  uint8_t nb_out[kMaxBlocks];    // N block -> dY tensor
  uint16_t nb_c0[kMaxBlocks];    // N block -> first channel inside that tensor
  uint8_t blk_src[kMaxBlocks];   // M block -> gathered tensor
  uint16_t blk_c0[kMaxBlocks];   // M block -> first channel inside that tensor
  uint16_t blk_goff[kMaxBlocks]; // M block -> first channel over the virtual concatenation
  uint8_t blk_tap[kMaxBlocks];   // taps-in-M mode: M block -> tap (layers with <= 64 gathered channels put
                                 // 128/ck different TAPS of the single channel block into one M tile)
  int8_t blk_ka[kMaxBlocks], blk_kb[kMaxBlocks], blk_kc[kMaxBlocks];   // ... and its (kd, kh, kw) index
  int taps_in_m;
  int nblocks;                   // total M blocks (ck channels each)
  int blocks_per_tile;           // 128 / ck
  int cin_total;
  int kd, kh, kw, pd, ph, pw;
  int sd, sh, sw;                // FWD stride: A boxes are loaded with TMA element strides
  int tpg;                       // taps per group (1 or kw)
  int shift;                     // SHIFT mode: the tpg taps read ONE X box at row shifts 0..tpg-1
  uint32_t a_box_bytes;          // bytes TMA writes per X block (a_blk_bytes minus the pad rows in SHIFT mode)
  int mpg;                       // M tiles (of 128 rows) per CTA: they share the dY tile of a stage
  int bd, bh, bw, td, th, tw;    // brick, bricks per dim
  int batch;
  int kv;                        // voxels per brick (multiple of 16)
  int ck, cb;                    // channels per A / B block
  int n_tile, n_blocks;          // N per CTA, B blocks per CTA
  int co;                        // produced channels over all fused dY tensors
  int nout;
  int out_start[M1_MAX_OUT + 1];
  int stages;
  uint32_t a_tap_bytes, b_off, stage_bytes;
  uint32_t a_blk_bytes, b_blk_bytes;
  uint32_t tmem_cols;
  uint32_t idesc;
  uint32_t a_desc_hi, b_desc_hi;
  uint32_t a_lbo, b_lbo;         // >> 4
  int splits;
  // Split-K through partial tiles instead of atomics. With atomics every split adds its 128 x N tile onto the SAME
  // addresses; for thin launches (one tile, ~300 splits) that same-address reduction traffic in L2 IS the kernel:
  // 96 us of a 102 us 128->128 1x1x1 weight gradient remain with the loads switched off (M1_WG_NOLOAD). Launches
  // whose partial tiles [split][base CTA][128][cols] fit the context scratch store them (plain coalesced stores)
  // and wgrad_reduce_kernel folds the splits (1-8 interleaved groups of them per output element).
  float* partial;
  int dbg_noload;                // experiment: producer arrives without loading (MMA-issue-rate probe)
  int64_t bricks_total;
  float* dw[M1_MAX_OUT];
  int64_t st[M1_MAX_OUT], sr[M1_MAX_OUT], so[M1_MAX_OUT];
};

__global__ void __launch_bounds__(kThreads) wgrad_tc_kernel(const __grid_constant__ WgParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_full = smem_base;
  const uint32_t bar_empty = smem_base + 8u * 16u;
  const uint32_t bar_accum = smem_base + 8u * 32u;
  const uint32_t tmem_slot = smem_base + 8u * 33u;
  const uint32_t tiles = smem_base + 1024u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5;

  // ---- work decomposition: blockIdx.x = ((group * m_tiles + m_tile) * n_tiles + n_tile), blockIdx.y = split
  const int m_sub = (p.nblocks + p.blocks_per_tile - 1) / p.blocks_per_tile;   // 128-row tiles in total
  const int m_tiles = (m_sub + p.mpg - 1) / p.mpg;
  const int n_tiles = (p.co + p.n_tile - 1) / p.n_tile;
  int t = blockIdx.x;
  const int nt = t % n_tiles; t /= n_tiles;
  const int mt = t % m_tiles; t /= m_tiles;
  const int group = t;
  const int groups_per_row = p.kw / p.tpg;
  const int kw0 = p.taps_in_m ? 0 : (group % groups_per_row) * p.tpg;
  const int kh_i = p.taps_in_m ? 0 : (group / groups_per_row) % p.kh;
  const int kd_i = p.taps_in_m ? 0 : group / (groups_per_row * p.kh);
  const int n0 = nt * p.n_tile;
  const int msub = min(p.mpg, m_sub - mt * p.mpg);       // 128-row tiles this CTA owns
  const int64_t per = (p.bricks_total + p.splits - 1) / p.splits;
  const int64_t b_begin = (int64_t)blockIdx.y * per;
  const int64_t b_end = min(b_begin + per, p.bricks_total);
  const int iters = (int)max((int64_t)0, b_end - b_begin);

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
  if (p.shift) {
    // pad rows behind every X block (read by the shifted descriptors of taps 1, 2; never written by TMA)
    const uint32_t pad = p.a_blk_bytes - p.a_box_bytes;                  // 8 rows
    const uint32_t blocks_per_stage = p.b_off / p.a_blk_bytes;
    for (int s = 0; s < p.stages; ++s)
      for (uint32_t b = 0; b < blocks_per_stage; ++b) {
        uint8_t* q = smem_gen + 1024u + s * p.stage_bytes + b * p.a_blk_bytes + p.a_box_bytes;
        for (uint32_t i = threadIdx.x * 16u; i < pad; i += kThreads * 16u)
          *reinterpret_cast<uint4*>(q + i) = make_uint4(0u, 0u, 0u, 0u);
      }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + 8u * 33u);

  if (iters > 0) {
    if (warp == 0) {
      if (elect_one()) {
        // ===== TMA producer: one stage = one voxel brick. 32-bit arithmetic and an incrementally advanced
        // brick coordinate: 64-bit div/mod per stage in this single thread used to cost more than the MMAs.
        uint32_t stage = 0, phase = 0;
        int bi = (int)b_begin;
        int tw_i = bi % p.tw; bi /= p.tw;
        int th_i = bi % p.th; bi /= p.th;
        int td_i = bi % p.td;
        int n_img = bi / p.td;
        int tot_blk = 0;
        for (int mi = 0; mi < msub; ++mi)
          tot_blk += min(p.blocks_per_tile, p.nblocks - (mt * p.mpg + mi) * p.blocks_per_tile);
        const int a_slots = p.shift ? 1 : p.tpg;             // X boxes per block: one per tap, or one shared
        const uint32_t tx = (uint32_t)(a_slots * tot_blk) * p.a_box_bytes + (uint32_t)p.n_blocks * p.b_blk_bytes;
        const int nb0 = n0 / p.cb;
        for (int it = 0; it < iters; ++it) {
          const int d0 = td_i * p.bd, h0 = th_i * p.bh, w0 = tw_i * p.bw;
          mbar_wait(bar_empty + 8u * stage, phase ^ 1u);
          const uint32_t full = bar_full + 8u * stage;
          if (p.dbg_noload) {
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(full) : "memory");
          } else {
            mbar_expect_tx(full, tx);
            const uint32_t sbase = tiles + stage * p.stage_bytes;
            // (SHIFT mode: tpg == kw, kw0 == 0 - the one shared X box sits at tap c = 0, one column left of dY)
            const int aw = w0 * p.sw - p.pw, ah = h0 * p.sh - p.ph, ad = d0 * p.sd - p.pd;
            uint32_t dst = sbase;
            for (int mi = 0; mi < msub; ++mi) {
              const int blk0 = (mt * p.mpg + mi) * p.blocks_per_tile;
              const int nblk = min(p.blocks_per_tile, p.nblocks - blk0);
              for (int tp = 0; tp < a_slots; ++tp) {
                uint32_t dj = dst;
                for (int j = 0; j < nblk; ++j) {
                  const int blk = blk0 + j;
                  // taps-in-M: the block's own tap (host table), else the tap of this CTA's group
                  const int a = p.taps_in_m ? (int)p.blk_ka[blk] : kd_i, b = p.taps_in_m ? (int)p.blk_kb[blk] : kh_i,
                            c = p.taps_in_m ? (int)p.blk_kc[blk] : kw0 + tp;
                  tma_load_5d(dj, &p.tmA[p.blk_src[blk]], full, (int)p.blk_c0[blk], aw + c, ah + b, ad + a, n_img);
                  dj += p.a_blk_bytes;
                }
                dst += p.a_tap_bytes;
              }
            }
            uint32_t db = sbase + p.b_off;
            for (int j = 0; j < p.n_blocks; ++j) {
              const int nb = nb0 + j;
              tma_load_5d(db, &p.tmB[p.nb_out[nb]], full, (int)p.nb_c0[nb], w0, h0, d0, n_img);
              db += p.b_blk_bytes;
            }
          }
          if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1u; }
          if (++tw_i == p.tw) {
            tw_i = 0;
            if (++th_i == p.th) {
              th_i = 0;
              if (++td_i == p.td) { td_i = 0; ++n_img; }
            }
          }
        }
      }
      __syncwarp();
    } else if (warp == 1) {
      if (elect_one()) {
        // ===== MMA issuer (32-bit descriptor arithmetic only: the issuing thread's instruction latency is exposed)
        uint32_t stage = 0, phase = 0;
        const uint64_t a_hi = (uint64_t)p.a_desc_hi << 32, b_hi = (uint64_t)p.b_desc_hi << 32;
        const uint32_t a_kstep = (16u * p.ck * 2u) >> 4, b_kstep = (16u * p.cb * 2u) >> 4;
        const uint32_t a_tap16 = p.a_tap_bytes >> 4;
        const uint32_t a_row16 = (2u * p.ck) >> 4;           // SHIFT mode: tap c reads the shared box c rows further
        const uint32_t a_lbo = p.a_lbo << 16, b_lbo = p.b_lbo << 16;
        const int k16s = p.kv / 16;
        for (int it = 0; it < iters; ++it) {
          mbar_wait(bar_full + 8u * stage, phase);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t sbase = tiles + stage * p.stage_bytes;
          const uint32_t b0 = ((sbase + p.b_off) >> 4) | b_lbo;
          uint32_t a_m = (sbase >> 4) | a_lbo;
          uint32_t d_t = tmem_base;
          for (int mi = 0; mi < msub; ++mi) {
            uint32_t a_t = a_m;
            for (int tp = 0; tp < p.tpg; ++tp) {   // accumulator = (M sub-tile, tap)
              uint32_t a_k = a_t, b_k = b0;
              umma_f16(d_t, a_hi | a_k, b_hi | b_k, p.idesc, it ? 1u : 0u);
              for (int k = 1; k < k16s; ++k) {
                a_k += a_kstep; b_k += b_kstep;
                umma_f16(d_t, a_hi | a_k, b_hi | b_k, p.idesc, 1u);
              }
              a_t += p.shift ? a_row16 : a_tap16;
              d_t += (uint32_t)p.n_tile;
            }
            a_m += p.shift ? a_tap16 : a_tap16 * (uint32_t)p.tpg;
          }
          umma_commit(bar_empty + 8u * stage);
          if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1u; }
        }
        umma_commit(bar_accum);
      }
      __syncwarp();
    }

    // ===== epilogue: TMEM -> registers -> fp32 atomics into dW =====
    mbar_wait(bar_accum, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int r = threadIdx.x;                       // row of the M tile == TMEM lane
    const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
    for (int mi = 0; mi < msub; ++mi)
    for (int tp = 0; tp < p.tpg; ++tp) {
      const int blk0 = (mt * p.mpg + mi) * p.blocks_per_tile;
      const int nblk = min(p.blocks_per_tile, p.nblocks - blk0);
      const int rblk = blk0 + r / p.ck;
      const bool row_ok = (r / p.ck) < nblk;
      const int rglob = row_ok ? (int)p.blk_goff[rblk] + r % p.ck : 0;
      const int tap = (p.taps_in_m && row_ok) ? (int)p.blk_tap[rblk] : (kd_i * p.kh + kh_i) * p.kw + kw0 + tp;
      for (int j = 0; j < p.n_tile; j += 8) {
        uint32_t v[8];
        tmem_ld8(lane_addr + (uint32_t)((mi * p.tpg + tp) * p.n_tile + j), v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (p.partial != nullptr) {
          const int cols = p.mpg * p.tpg * p.n_tile;
          float* q = p.partial + (((int64_t)blockIdx.y * gridDim.x + blockIdx.x) * 128 + r) * cols +
                     (mi * p.tpg + tp) * p.n_tile + j;
          *reinterpret_cast<uint4*>(q) = make_uint4(v[0], v[1], v[2], v[3]);
          *reinterpret_cast<uint4*>(q + 4) = make_uint4(v[4], v[5], v[6], v[7]);
          continue;
        }
        if (!row_ok) continue;
        int n = n0 + j;                       // 8-column groups never straddle two dY tensors (channels % 16 == 0)
        if (n >= p.co) continue;
        int o = 0;
        while (o + 1 < p.nout && n >= p.out_start[o + 1]) ++o;
        n -= p.out_start[o];
        float* dst = p.dw[o] + tap * p.st[o] + (int64_t)rglob * p.sr[o] + (int64_t)n * p.so[o];
#pragma unroll
        for (int i = 0; i < 8; ++i) atomicAdd(dst + (int64_t)i * p.so[o], __uint_as_float(v[i]));
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols)
                 : "memory");
  }
}

// dW += the partial tiles of the splits. grid = (base CTAs, chunks of (row, 4-column) elements, split groups): a
// thread sums every gridDim.z-th split of its element and adds the result with ONE atomic per value - gridDim.z-way
// instead of splits-way contention. The address logic mirrors the epilogue of wgrad_tc_kernel.
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const __grid_constant__ WgParams p, int base_ctas) {
  const int cols = p.mpg * p.tpg * p.n_tile, c4 = cols / 4;
  const int idx = blockIdx.y * 256 + threadIdx.x;
  if (idx >= 128 * c4) return;
  const int r = idx / c4, c = (idx % c4) * 4;
  const int m_sub = (p.nblocks + p.blocks_per_tile - 1) / p.blocks_per_tile;
  const int m_tiles = (m_sub + p.mpg - 1) / p.mpg;
  const int n_tiles = (p.co + p.n_tile - 1) / p.n_tile;
  int t = blockIdx.x;
  const int nt = t % n_tiles; t /= n_tiles;
  const int mt = t % m_tiles; t /= m_tiles;
  const int group = t;
  const int groups_per_row = p.kw / p.tpg;
  const int kw0 = p.taps_in_m ? 0 : (group % groups_per_row) * p.tpg;
  const int kh_i = p.taps_in_m ? 0 : (group / groups_per_row) % p.kh;
  const int kd_i = p.taps_in_m ? 0 : group / (groups_per_row * p.kh);
  const int acc = c / p.n_tile, j = c % p.n_tile, mi = acc / p.tpg, tp = acc % p.tpg;
  if (mi >= min(p.mpg, m_sub - mt * p.mpg)) return;
  const int blk0 = (mt * p.mpg + mi) * p.blocks_per_tile;
  const int nblk = min(p.blocks_per_tile, p.nblocks - blk0);
  if (r / p.ck >= nblk) return;
  const int rblk = blk0 + r / p.ck;
  const int rglob = (int)p.blk_goff[rblk] + r % p.ck;
  const int tap = p.taps_in_m ? (int)p.blk_tap[rblk] : (kd_i * p.kh + kh_i) * p.kw + kw0 + tp;
  int n = nt * p.n_tile + j;
  if (n >= p.co) return;
  int o = 0;
  while (o + 1 < p.nout && n >= p.out_start[o + 1]) ++o;
  n -= p.out_start[o];
  const int64_t per = (p.bricks_total + p.splits - 1) / p.splits;
  const int nsplit = (int)min((int64_t)p.splits, (p.bricks_total + per - 1) / per);     // splits that own bricks
  const float* q = p.partial + ((int64_t)blockIdx.x * 128 + r) * cols + c;
  const int64_t stride = (int64_t)base_ctas * 128 * cols;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
  for (int sp = blockIdx.z; sp < nsplit; sp += gridDim.z) {
    const float4 v = __ldcg(reinterpret_cast<const float4*>(q + sp * stride));
    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
  }
  float* dst = p.dw[o] + tap * p.st[o] + (int64_t)rglob * p.sr[o] + (int64_t)n * p.so[o];
  atomicAdd(dst, s.x);
  atomicAdd(dst + p.so[o], s.y);
  atomicAdd(dst + 2 * p.so[o], s.z);
  atomicAdd(dst + 3 * p.so[o], s.w);
}

struct WgPlan {
  int taps_in_m, mpg, shift;
  uint32_t a_box_bytes;
  int ck, cb, n_tile, n_blocks, tpg, kv, bd, bh, bw, td, th, tw, stages, nblocks, cin_total;
  uint32_t a_blk_bytes, b_blk_bytes, a_tap_bytes, b_off, stage_bytes, smem_bytes, tmem_cols;
};

bool make_wg_plan(const m1_conv_desc* d, int j0, int jn, WgPlan* pl) {
  if (d->mode != M1_CONV_FWD) return false;
  // activations and output gradients must share one 16-bit format (kind::f16 with mixed operand formats traps)
  if (!m1_is16(d->act_dtype) || d->out_dtype != d->act_dtype) return false;
  for (int i = 0; i < 3; ++i)
    if (d->stride[i] < 1 || d->stride[i] > 2) return false;
  if (d->nsrc < 1 || d->nsrc > M1_MAX_SRC) return false;
  int ck = 64, cin = 0;
  for (int s = 0; s < d->nsrc; ++s) {
    if (d->src_c[s] % 16) return false;
    while (d->src_c[s] % ck) ck >>= 1;
    cin += d->src_c[s];
  }
  if (cin / ck > kMaxBlocks) {
    // fall back to coarser blocks only if every tensor allows it; otherwise refuse
    return false;
  }
  int co = 0, cb = 64;
  for (int j = j0; j < j0 + jn; ++j) {
    if (d->out_c[j] % 16) return false;
    while (d->out_c[j] % cb) cb >>= 1;
    co += d->out_c[j];
  }
  if (co / cb > kMaxBlocks) return false;
  int n_tile = co;
  if (n_tile > 256) {
    // largest N tile <= 256 that is a multiple of the block size and divides the total
    n_tile = 0;
    for (int c = 256 / cb * cb; c >= cb; c -= cb)
      if (co % c == 0) { n_tile = c; break; }
    if (!n_tile) return false;
  }
  int tpg = (d->kernel[2] * n_tile <= 512) ? d->kernel[2] : 1;
  const int ntaps = d->kernel[0] * d->kernel[1] * d->kernel[2];
  // few gathered channels (one block per tap): fill the 128 M rows with 128/ck different taps
  const bool taps_in_m = (cin == ck) && ck < 128 && ntaps > 1 && ntaps <= kMaxBlocks;
  if (taps_in_m) tpg = 1;
  // M tiles per CTA (they share the dY tile): tune[3], default 1
  const int m_sub_total = ((taps_in_m ? ntaps : cin / ck) + 128 / ck - 1) / (128 / ck);
  int mpg_req = d->tune[3] > 1 ? d->tune[3] : (getenv("M1_WG_MPG") ? atoi(getenv("M1_WG_MPG")) : 1);
  if (mpg_req > m_sub_total) mpg_req = m_sub_total;
  if (mpg_req < 1) mpg_req = 1;
  // brick: voxels multiple of 16, <= kv_max, best volume coverage
  const int D = d->out_dhw[0], H = d->out_dhw[1], W = d->out_dhw[2];
  auto pick = [&](int kv_max, int* obd, int* obh, int* obw) {
    double best = -1;
    for (int bd = 1; bd <= kv_max && bd <= D; ++bd)
      for (int bh = 1; bd * bh <= kv_max && bh <= H; ++bh)
        for (int bw = 1; bd * bh * bw <= kv_max && bw <= W; ++bw) {
          const int kv = bd * bh * bw;
          if (kv % 16) continue;
          if (bd * d->stride[0] > 256 || bh * d->stride[1] > 256 || bw * d->stride[2] > 256) continue;
          const int64_t tiles = (int64_t)((D + bd - 1) / bd) * ((H + bh - 1) / bh) * ((W + bw - 1) / bw);
          const double eff = (double)D * H * W / ((double)kv * tiles);
          const double score = eff + 1e-3 * kv / kv_max + 1e-6 * bw;
          if (score > best) { best = score; *obd = bd; *obh = bh; *obw = bw; }
        }
    return best > 0;
  };
  // tiling overrides: host autotuning (d->tune) or environment (experiments)
  static const int g_kv = getenv("M1_WG_KV") ? atoi(getenv("M1_WG_KV")) : 0;
  static const int g_tpg = getenv("M1_WG_TPG") ? atoi(getenv("M1_WG_TPG")) : 0;
  static const int g_stages = getenv("M1_WG_STAGES") ? atoi(getenv("M1_WG_STAGES")) : 0;
  const int env_kv = d->tune[0] ? d->tune[0] : g_kv;
  const int env_tpg = d->tune[1] ? d->tune[1] : g_tpg;
  const int env_stages = d->tune[2] ? d->tune[2] : (g_stages ? g_stages : 2);
  if (env_tpg == 1) tpg = 1;
  pl->shift = 0;
  if (env_tpg == 2) {
    // ---- SHIFT mode: bh full-width lines at row pitch P >= W + 2 for both operands, bh * P % 16 == 0
    if (taps_in_m || d->kernel[2] != 3 || d->pad[2] != 1 || tpg != 3) return false;
    for (int i = 0; i < 3; ++i)
      if (d->stride[i] != 1) return false;
    const int kv_cap = env_kv ? std::max(env_kv, 48) : 192;
    double best = -1;
    int bP = 0, bbh = 0;
    for (int P = W + 2; P <= std::min(256, W + 2 + 15); ++P)
      for (int bh = 1; bh <= H && bh * P <= kv_cap; ++bh) {
        if ((bh * P) % 16) continue;
        const int th = (H + bh - 1) / bh;
        const double eff = (double)W * H / ((double)P * bh * th);          // useful K rows / issued K rows
        const double score = eff + 1e-3 * (bh * P) / kv_cap;
        if (score > best) { best = score; bP = P; bbh = bh; }
      }
    if (best < 0) return false;
    const int kv = bbh * bP;
    int mpg = mpg_req;
    while (mpg > 1 && mpg * tpg * n_tile > 512) --mpg;
    const uint32_t a_box = (uint32_t)kv * ck * 2u, a_blk = a_box + 8u * ck * 2u, b_blk = (uint32_t)kv * cb * 2u;
    const uint32_t a_tap = (128u / ck) * a_blk;
    const uint32_t b_off = (uint32_t)mpg * a_tap;
    const uint32_t stage = (b_off + (uint32_t)(n_tile / cb) * b_blk + 1023u) & ~1023u;
    int stages = (int)((227u * 1024u - 2048u) / stage);
    const int cap = d->tune[2] ? d->tune[2] : (g_stages ? g_stages : 4);
    if (stages > cap) stages = cap;
    if (stages < 2) return false;
    pl->shift = 1; pl->a_box_bytes = a_box;
    pl->ck = ck; pl->cb = cb; pl->n_tile = n_tile; pl->n_blocks = n_tile / cb; pl->tpg = tpg; pl->kv = kv;
    pl->bd = 1; pl->bh = bbh; pl->bw = bP;
    pl->td = D; pl->th = (H + bbh - 1) / bbh; pl->tw = 1;
    pl->stages = stages; pl->nblocks = cin / ck; pl->cin_total = cin;
    pl->taps_in_m = 0; pl->mpg = mpg;
    pl->a_blk_bytes = a_blk; pl->b_blk_bytes = b_blk; pl->a_tap_bytes = a_tap; pl->b_off = b_off;
    pl->stage_bytes = stage; pl->smem_bytes = 2048u + (uint32_t)stages * stage;
    uint32_t cols = 32;
    while ((int)cols < tpg * mpg * n_tile) cols <<= 1;
    pl->tmem_cols = cols;
    return true;
  }
  for (int kv_max = env_kv ? env_kv : 128; kv_max >= 16; kv_max >>= 1) {
    int bd = 1, bh = 1, bw = 1;
    if (!pick(kv_max, &bd, &bh, &bw)) continue;
    const int kv = bd * bh * bw;
    for (int tp = tpg; tp >= 1; tp = (tp == 1 ? 0 : 1)) {
      int mpg = mpg_req;
      while (mpg > 1 && mpg * tp * n_tile > 512) --mpg;
      const uint32_t a_blk = (uint32_t)kv * ck * 2u, b_blk = (uint32_t)kv * cb * 2u;
      const uint32_t a_tap = (128u / ck) * a_blk;
      const uint32_t b_off = (uint32_t)(tp * mpg) * a_tap;
      const uint32_t stage = (b_off + (uint32_t)(n_tile / cb) * b_blk + 1023u) & ~1023u;
      int stages = (int)((227u * 1024u - 2048u) / stage);
      if (stages > 8) stages = 8;
      if (env_stages && stages > env_stages) stages = env_stages;
      if (stages < 2) continue;
      pl->ck = ck; pl->cb = cb; pl->n_tile = n_tile; pl->n_blocks = n_tile / cb; pl->tpg = tp; pl->kv = kv;
      pl->bd = bd; pl->bh = bh; pl->bw = bw;
      pl->td = (D + bd - 1) / bd; pl->th = (H + bh - 1) / bh; pl->tw = (W + bw - 1) / bw;
      pl->stages = stages; pl->nblocks = taps_in_m ? ntaps : cin / ck; pl->cin_total = cin;
      pl->taps_in_m = taps_in_m ? 1 : 0;
      pl->mpg = mpg;
      pl->a_box_bytes = a_blk;
      pl->a_blk_bytes = a_blk; pl->b_blk_bytes = b_blk; pl->a_tap_bytes = a_tap; pl->b_off = b_off;
      pl->stage_bytes = stage; pl->smem_bytes = 2048u + (uint32_t)stages * stage;
      uint32_t cols = 32;
      while ((int)cols < tp * mpg * n_tile) cols <<= 1;
      pl->tmem_cols = cols;
      return true;
    }
  }
  return false;
}

}  // namespace

int m1_conv3d_wgrad_tc_supported(const m1_conv_desc* d, int j0, int jn) {
  WgPlan pl;
  return make_wg_plan(d, j0, jn, &pl) ? 1 : 0;
}

// out: ck, cb, n_tile, tpg, mpg, kv, bd, bh, bw, stages, smem_bytes, tmem_cols, shift, taps_in_m, stage_bytes
int m1_conv3d_wgrad_plan_info(const m1_conv_desc* d, int32_t* out) {
  WgPlan pl;
  if (!make_wg_plan(d, 0, d->nout, &pl) && !make_wg_plan(d, 0, 1, &pl)) return 0;
  const int32_t v[15] = {pl.ck, pl.cb, pl.n_tile, pl.tpg, pl.mpg, pl.kv, pl.bd, pl.bh, pl.bw, pl.stages,
                         (int32_t)pl.smem_bytes, (int32_t)pl.tmem_cols, pl.shift, pl.taps_in_m,
                         (int32_t)pl.stage_bytes};
  for (int i = 0; i < 15; ++i) out[i] = v[i];
  return 15;
}

extern "C" int m1_conv3d_wgrad_tc_supported0(const m1_conv_desc* d) {
  return m1_conv3d_wgrad_tc_supported(d, 0, d->nout) || m1_conv3d_wgrad_tc_supported(d, 0, 1);
}

int m1_conv3d_wgrad_tc(m1_ctx* ctx, const m1_conv_desc* d, int j0, int jn, const void* const* srcs,
                       const void* const* douts, float* const* dws, cudaStream_t st) {
  WgPlan pl;
  M1_CHECK(make_wg_plan(d, j0, jn, &pl), "m1_conv3d_wgrad: launch not supported by the tcgen05 engine");
  M1_CHECK(ctx->encode_tiled != nullptr, "cuTensorMapEncodeTiled entry point unavailable");
  EncodeTiledFn encode = reinterpret_cast<EncodeTiledFn>(ctx->encode_tiled);
  static_assert(sizeof(WgParams) < 4000, "kernel parameter block too large");
  WgParams p;
  memset(&p, 0, sizeof(p));
  const int D = d->out_dhw[0], H = d->out_dhw[1], W = d->out_dhw[2];
  int blk = 0, goff = 0;
  for (int s = 0; s < d->nsrc; ++s) {
    M1_CHECK(((uintptr_t)srcs[s] & 15) == 0, "m1_conv3d_wgrad: gathered tensor %d not 16-byte aligned", s);
    int r = encode_ndhwc(encode, &p.tmA[s], srcs[s], d->act_dtype, d->src_c[s], d->in_dhw[2], d->in_dhw[1], d->in_dhw[0],
                         d->batch, pl.ck, pl.bw, pl.bh, pl.bd, d->stride[2], d->stride[1], d->stride[0]);
    M1_CHECK(r == 0, "cuTensorMapEncodeTiled(wgrad A %d) failed: %d", s, r);
    if (!pl.taps_in_m) {
      for (int c0 = 0; c0 < d->src_c[s]; c0 += pl.ck) {
        p.blk_src[blk] = (uint8_t)s;
        p.blk_c0[blk] = (uint16_t)c0;
        p.blk_goff[blk] = (uint16_t)(goff + c0);
        ++blk;
      }
    }
    goff += d->src_c[s];
  }
  if (pl.taps_in_m) {
    // one channel block (of gathered tensor 0, channels [0, ck)) per tap
    const int ntaps = d->kernel[0] * d->kernel[1] * d->kernel[2];
    for (int t = 0; t < ntaps; ++t) {
      p.blk_src[t] = 0; p.blk_c0[t] = 0; p.blk_goff[t] = 0; p.blk_tap[t] = (uint8_t)t;
      p.blk_kc[t] = (int8_t)(t % d->kernel[2]);
      p.blk_kb[t] = (int8_t)((t / d->kernel[2]) % d->kernel[1]);
      p.blk_ka[t] = (int8_t)(t / (d->kernel[2] * d->kernel[1]));
    }
  }
  p.taps_in_m = pl.taps_in_m;
  {
    int nb = 0, acc = 0;
    for (int o = 0; o < jn; ++o) {
      const int j = j0 + o;
      int r = encode_ndhwc(encode, &p.tmB[o], douts[j], d->out_dtype, d->out_c[j], W, H, D, d->batch, pl.cb, pl.bw, pl.bh,
                           pl.bd);
      M1_CHECK(r == 0, "cuTensorMapEncodeTiled(wgrad B %d) failed: %d", j, r);
      for (int c0 = 0; c0 < d->out_c[j]; c0 += pl.cb) {
        p.nb_out[nb] = (uint8_t)o;
        p.nb_c0[nb] = (uint16_t)c0;
        ++nb;
      }
      p.out_start[o] = acc;
      acc += d->out_c[j];
      p.dw[o] = dws[j];
      p.st[o] = d->w_stride_tap[j]; p.sr[o] = d->w_stride_red[j]; p.so[o] = d->w_stride_out[j];
    }
    p.out_start[jn] = acc;
    p.nout = jn;
    p.co = acc;
  }
  p.nsrc = d->nsrc;
  p.nblocks = pl.nblocks;
  p.blocks_per_tile = 128 / pl.ck;
  p.cin_total = pl.cin_total;
  p.kd = d->kernel[0]; p.kh = d->kernel[1]; p.kw = d->kernel[2];
  p.pd = d->pad[0]; p.ph = d->pad[1]; p.pw = d->pad[2];
  p.sd = d->stride[0]; p.sh = d->stride[1]; p.sw = d->stride[2];
  p.tpg = pl.tpg;
  p.mpg = pl.mpg;
  p.shift = pl.shift;
  p.a_box_bytes = pl.a_box_bytes;
  p.bd = pl.bd; p.bh = pl.bh; p.bw = pl.bw; p.td = pl.td; p.th = pl.th; p.tw = pl.tw;
  p.batch = d->batch;
  p.kv = pl.kv;
  p.ck = pl.ck; p.cb = pl.cb;
  p.n_tile = pl.n_tile; p.n_blocks = pl.n_blocks;
  p.stages = pl.stages;
  p.a_tap_bytes = pl.a_tap_bytes; p.b_off = pl.b_off; p.stage_bytes = pl.stage_bytes;
  p.a_blk_bytes = pl.a_blk_bytes; p.b_blk_bytes = pl.b_blk_bytes;
  p.tmem_cols = pl.tmem_cols;
  // D = f32, A = activations, B = output gradients (one common format: bf16, or f16), both MN-major (bits 15, 16),
  // N >> 3 at [17,23), M >> 4 at [24,29)
  p.idesc = (1u << 4) | (idesc_fmt(d->act_dtype) << 7) | (idesc_fmt(d->out_dtype) << 10) | (1u << 15) | (1u << 16) |
            ((uint32_t)(pl.n_tile >> 3) << 17) | ((128u >> 4) << 24);
  // MN-major descriptors: SBO = 8 voxel rows of a block, LBO = one block (next channel block)
  p.a_desc_hi = ((8u * pl.ck * 2u) >> 4) | (1u << 14) | (layout_for(pl.ck) << 29);
  p.b_desc_hi = ((8u * pl.cb * 2u) >> 4) | (1u << 14) | (layout_for(pl.cb) << 29);
  p.a_lbo = pl.a_blk_bytes >> 4;
  p.b_lbo = pl.b_blk_bytes >> 4;
  M1_CHECK(p.a_lbo < (1u << 14) && p.b_lbo < (1u << 14), "wgrad: block too large for the descriptor LBO field");
  p.bricks_total = (int64_t)d->batch * pl.td * pl.th * pl.tw;
  const int m_tiles = ((pl.nblocks + p.blocks_per_tile - 1) / p.blocks_per_tile + pl.mpg - 1) / pl.mpg;
  const int n_tiles = (p.co + pl.n_tile - 1) / pl.n_tile;
  const int groups = pl.taps_in_m ? 1 : p.kd * p.kh * (p.kw / pl.tpg);
  const int64_t base_ctas = (int64_t)groups * m_tiles * n_tiles;
  // split-K factor: ~2-4 CTAs per SM in total, chosen so that the last wave is as full as possible
  int64_t splits = 1;
  {
    const int64_t sms = ctx->num_sms, max_splits = std::max<int64_t>(1, p.bricks_total / 4);
    double best = -1;
    for (int64_t sp = 1; sp <= max_splits && base_ctas * sp <= sms * 6; ++sp) {
      const int64_t total = base_ctas * sp;
      const double waves = (double)total / sms;
      const double eff = waves / std::ceil(waves);          // fill of the last wave
      const double score = eff - (total < sms * 2 ? 0.5 * (1.0 - (double)total / (sms * 2)) : 0.0) - 1e-3 * sp;
      if (score > best) { best = score; splits = sp; }
    }
  }
  p.splits = (int)splits;
  {
    static const int g_scr = getenv("M1_WG_SCRATCH") ? atoi(getenv("M1_WG_SCRATCH")) : 2;
    const size_t bytes = (size_t)splits * base_ctas * 128 * ((size_t)pl.mpg * pl.tpg * pl.n_tile) * sizeof(float);
    // 1: thin launches only (<= 8 tiles, >= 8 splits); 2 (default): every split-K launch whose partial tiles fit the
    // scratch - the transposed-convolution weight gradients (27 tiles x 10 splits) gain as much as the thin ones:
    // 0.195 -> 0.144 ms; whole step 74.96 -> 73.95 ms in one A/B call
    const bool thin = base_ctas <= 8 && splits >= 8;
    p.partial = (g_scr && bytes <= ctx->partial_bytes && (thin || (g_scr >= 2 && splits >= 2))) ? ctx->partial : nullptr;
  }
  static const int g_noload = getenv("M1_WG_NOLOAD") ? atoi(getenv("M1_WG_NOLOAD")) : 0;
  p.dbg_noload = g_noload;

  static int smem_set = 0;
  if (!smem_set) {
    M1_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    smem_set = 1;
  }
  dim3 grid((unsigned)base_ctas, (unsigned)splits);
  wgrad_tc_kernel<<<grid, kThreads, pl.smem_bytes, st>>>(p);
  M1_LAUNCH_CHECK(ctx);
  if (p.partial != nullptr) {
    const int chunks = (128 * (pl.mpg * pl.tpg * pl.n_tile / 4) + 255) / 256;
    // split groups per output element: enough threads to fill the machine (>= ~64 K), at most 8-way atomics
    const int64_t threads1 = (int64_t)base_ctas * 128 * (pl.mpg * pl.tpg * pl.n_tile / 4);
    const int groups_z = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(8, splits), (65536 + threads1 - 1) / threads1));
    wgrad_reduce_kernel<<<dim3((unsigned)base_ctas, (unsigned)chunks, (unsigned)groups_z), 256, 0, st>>>(p, (int)base_ctas);
    M1_LAUNCH_CHECK(ctx);
  }
  return 0;
}
