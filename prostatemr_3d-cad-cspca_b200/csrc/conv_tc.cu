// conv_tc.cu — K1/K2: 3-D convolution, transposed convolution and their data gradients as an
// implicit GEMM on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM), operands
// staged by TMA (cp.async.bulk.tensor) straight from the NDHWC activation tensors.
//
// Replaces tf.keras.layers.Conv3D at R:network_blocks.py:37-46 (conv1||conv4, conv2, conv3 of
// every SEResNetBottleNeck), R:networks.py:472, tf.keras.layers.Conv3DTranspose at
// R:networks.py:496-553 and the Conv3DBackpropInputV2 of their autodiff.
//
// Two gather forms cover all four cases (m1_conv_mode):
//   FWD         in = o*s + k - pad : the A box of a tap is loaded with TMA element strides s
//               (every s-th voxel), its corner at o0*s + k - pad
//   TRANSPOSED  in = (o + pad - k)/s : decomposed into the s_d*s_h*s_w output PHASES o = s*j + phi;
//               inside a phase it is a stride-1 gather in = j + (phi + pad - k)/s over the taps with
//               k == phi + pad (mod s); the tile is a brick of j, written back with stride s
//
// GEMM view   D[M = 128 output voxels, N = produced channels] += A[M, K] * B[N, K]^T
//   M tile  = a (bd x bh x bw) brick of output voxels of ONE volume (<= 128 voxels)
//   K       = taps x gathered channels, walked as "k-steps" of ck in {16,32,64} channels of one
//             tap of one gathered tensor (the channel concatenation of the reference is never
//             materialised: every gathered tensor has its own tensor map)
//   A tile  = one 5-D TMA box (ck, bw, bh, bd, 1) whose corner is shifted by the tap offset;
//             out-of-volume voxels are zero-filled by TMA == TF "SAME" padding
//   B tile  = one 3-D TMA box (ck, n_tile, 1) of the bf16 K-major weight pack [tap][N][K]
// Both land in shared memory in the canonical K-major swizzled UMMA layout (rows of ck*2 bytes,
// 8-row swizzle atoms), so one tcgen05.mma per 16 channels consumes them with no data movement
// by threads.  One CTA = one M tile x one N tile; several CTAs are co-resident per SM so that the
// epilogue of one overlaps the main loop of another. Warp 0 = elected TMA producer, warp 1 = elected MMA
// issuer (elect.sync, see tc_common.cuh), all 8 warps run the epilogue (tc_epilogue.cuh). Stride-1 launches
// with in-plane taps may be routed to the halo variant (conv_tc_halo.cu) - m1_conv_desc.tune[0].
#include "tc_epilogue.cuh"
#include <algorithm>
#include <cstdlib>
#include <vector>

namespace {
using namespace tc;

constexpr int kThreads = 256;   // warp 0: TMA producer, warp 1: MMA issuer, all 8 warps: epilogue

struct TcParams {
  CUtensorMap tmA[M1_MAX_SRC];
  CUtensorMap tmB;
  int nsrc;
  int src_chunks[M1_MAX_SRC];  // channels / ck of each gathered tensor
  int src_koff[M1_MAX_SRC];    // first K index of the tensor inside a weight-pack row
  // tap table: entries [phase_tap0[ph], phase_tap0[ph+1]) belong to output phase ph
  int nphase;
  uint8_t phase_tap0[9];
  int8_t phase_d[8], phase_h[8], phase_w[8];   // phi per dim
  uint8_t tap_w[32];                           // weight tap of the entry
  int8_t tap_od[32], tap_oh[32], tap_ow[32];   // gather offset of the entry (voxels of the gathered grid)
  int istr_d, istr_h, istr_w;  // gathered-grid step per tile voxel (FWD stride; 1 for phases)
  int ostr_d, ostr_h, ostr_w;  // produced-grid step per tile voxel (phase stride; 1 for FWD)
  int bd, bh, bw;              // brick (in tile voxels)
  int td, th, tw;              // bricks per dim
  int Do, Ho, Wo;              // produced grid
  int n_tile;
  int ck;
  int group, stages;
  uint32_t a_alloc, slot_bytes;
  uint32_t tx_per_kstep;
  uint32_t tmem_cols;
  uint32_t idesc;
  uint32_t desc_hi;            // upper 32 bits of the UMMA shared-memory descriptors
  EpiOut epi;
};

__global__ void __launch_bounds__(kThreads)
conv_tc_kernel(const __grid_constant__ TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  // barriers + TMEM slot live in the first 1 KiB after alignment; tiles follow, 1 KiB aligned
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_full = smem_base;                 // stages x 8 B
  const uint32_t bar_empty = smem_base + 8u * 16u;     // stages x 8 B (<= 16 stages)
  const uint32_t bar_accum = smem_base + 8u * 32u;
  const uint32_t tmem_slot = smem_base + 8u * 33u;
  const uint32_t tiles = smem_base + 1024u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  const int warp = threadIdx.x >> 5;

  // ---- tile coordinates -----------------------------------------------------------------
  int t = blockIdx.x;
  const int tw_i = t % p.tw; t /= p.tw;
  const int th_i = t % p.th; t /= p.th;
  const int td_i = t % p.td; t /= p.td;
  const int n_img = t;
  const int d0 = td_i * p.bd, h0 = th_i * p.bh, w0 = tw_i * p.bw;
  const int n0 = blockIdx.y * p.n_tile;
  const int ph = blockIdx.z;
  const int tap_begin = p.phase_tap0[ph], tap_end = p.phase_tap0[ph + 1];

  // ---- one-time setup -------------------------------------------------------------------
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     tmem_slot),
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
  const int ksteps = (tap_end - tap_begin) * total_chunks;
  const int nstage_iters = (ksteps + p.group - 1) / p.group;

  if (warp == 0) {
    if (elect_one()) {
      // ===== TMA producer =====
      int tap = tap_begin, src = 0, chunk = 0;
      const int a_d0 = d0 * p.istr_d, a_h0 = h0 * p.istr_h, a_w0 = w0 * p.istr_w;
      uint32_t stage = 0, phase = 0;
      for (int it = 0; it < nstage_iters; ++it) {
        const int g = min(p.group, ksteps - it * p.group);
        mbar_wait(bar_empty + 8u * stage, phase ^ 1u);
        const uint32_t full = bar_full + 8u * stage;
        mbar_expect_tx(full, p.tx_per_kstep * (uint32_t)g);
        for (int j = 0; j < g; ++j) {
          const uint32_t slot = tiles + (stage * p.group + j) * p.slot_bytes;
          const int c_in = chunk * p.ck;
          tma_load_5d(slot, &p.tmA[src], full, c_in, a_w0 + p.tap_ow[tap], a_h0 + p.tap_oh[tap],
                      a_d0 + p.tap_od[tap], n_img);
          tma_load_3d(slot + p.a_alloc, &p.tmB, full, p.src_koff[src] + c_in, n0, (int)p.tap_w[tap]);
          // advance (chunk, src, tap)
          if (++chunk == p.src_chunks[src]) {
            chunk = 0;
            if (++src == p.nsrc) { src = 0; ++tap; }
          }
        }
        if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1u; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (elect_one()) {
      // ===== MMA issuer =====
      uint32_t stage = 0, phase = 0;
      const uint64_t hi = (uint64_t)p.desc_hi << 32;
      const int k16s = p.ck / 16;
      uint32_t acc = 0;
      for (int it = 0; it < nstage_iters; ++it) {
        const int g = min(p.group, ksteps - it * p.group);
        mbar_wait(bar_full + 8u * stage, phase);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int j = 0; j < g; ++j) {
          const uint32_t slot = tiles + (stage * p.group + j) * p.slot_bytes;
          const uint64_t a_lo = (uint64_t)((slot >> 4) & 0x3FFFu);
          const uint64_t b_lo = (uint64_t)(((slot + p.a_alloc) >> 4) & 0x3FFFu);
          for (int k = 0; k < k16s; ++k) {
            // 16 channels = 32 bytes further along the swizzled row: +2 in 16-byte units
            umma_f16(tmem_base, hi | (a_lo + 2u * k), hi | (b_lo + 2u * k), p.idesc, acc);
            acc = 1;
          }
        }
        umma_commit(bar_empty + 8u * stage);
        if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1u; }
      }
      umma_commit(bar_accum);
    }
    __syncwarp();
  }

  // ===== epilogue: 8 warps; warps w and w+4 share TMEM lanes 32*(w%4).. and split the columns =====
  // (a phase without taps - kernel smaller than the stride - produces bias only: nothing was issued,
  //  the MMA warp still commits bar_accum and the accumulator is treated as zero)
  mbar_wait(bar_accum, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  {
    const int r = (warp & 3) * 32 + (threadIdx.x & 31);  // row of the tile == TMEM lane
    const int lw = r % p.bw;
    const int lh = (r / p.bw) % p.bh;
    const int ld = r / (p.bw * p.bh);
    const int d = (d0 + ld) * p.ostr_d + p.phase_d[ph], h = (h0 + lh) * p.ostr_h + p.phase_h[ph],
              w = (w0 + lw) * p.ostr_w + p.phase_w[ph];
    const bool valid = (ld < p.bd) && d < p.Do && h < p.Ho && w < p.Wo;
    const int64_t vox = (((int64_t)n_img * p.Do + d) * p.Ho + h) * p.Wo + w;
    const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    int cb, ce;
    epi_cols(warp, p.n_tile, &cb, &ce);
    epilogue_row(p.epi, lane_addr, n0, cb, ce, valid, vox, ksteps == 0);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(p.tmem_cols)
                 : "memory");
  }
}

// ---- multi-tile variant (m1_conv_desc.tune[0] == 3: one of the variants the host's one-off autotuning times)
// Launches with a handful of k-steps per tile (1x1x1 convolutions, the output phases of transposed
// convolutions, few-channel layers: ~24 ms of the step at 20-400 TFLOP/s) are dominated by per-CTA fixed costs:
// TMEM allocation, barrier initialisation, the first TMA round trip, the drain of the epilogue. Here a CTA owns
// `tiles_per_cta` consecutive M tiles: the producer warp streams k-steps across tile boundaries (the smem ring
// never drains), accumulators are double-buffered in TMEM (tile i+1 is multiplied while tile i is stored), and
// eight dedicated epilogue warps (warps 2-9; warp w reads TMEM lanes 32*(w%4).., warps 2-5 / 6-9 split the
// columns) hand each accumulator back through bar_free.
constexpr int kThreadsMulti = 320;

__global__ void __launch_bounds__(kThreadsMulti)
conv_tc_multi_kernel(const __grid_constant__ TcParams p, int tiles_per_cta, int total_tiles, uint32_t acc_cols) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_full = smem_base;                 // stages x 8 B
  const uint32_t bar_empty = smem_base + 8u * 16u;     // stages x 8 B (<= 16 stages)
  const uint32_t bar_accum = smem_base + 8u * 32u;     // 2 x 8 B: accumulator buffer complete (MMA -> epilogue)
  const uint32_t bar_free = smem_base + 8u * 34u;      // 2 x 8 B: accumulator buffer drained (epilogue -> MMA)
  const uint32_t tmem_slot = smem_base + 8u * 36u;
  const uint32_t tiles = smem_base + 1024u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5;

  const int t_begin = blockIdx.x * tiles_per_cta;
  const int t_end = min(t_begin + tiles_per_cta, total_tiles);
  const int n0 = blockIdx.y * p.n_tile;
  const int ph = blockIdx.z;
  const int tap_begin = p.phase_tap0[ph], tap_end = p.phase_tap0[ph + 1];

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                 "r"(2u * acc_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 32) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(bar_full + 8u * s, 1);
      mbar_init(bar_empty + 8u * s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_accum + 8u * b, 1);
      mbar_init(bar_free + 8u * b, 8);                 // one arrival per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + 8u * 36u);

  int total_chunks = 0;
  for (int s = 0; s < p.nsrc; ++s) total_chunks += p.src_chunks[s];
  const int ksteps = (tap_end - tap_begin) * total_chunks;
  const int nstage_iters = (ksteps + p.group - 1) / p.group;

  if (warp == 0) {
    if (elect_one()) {
      // ===== TMA producer: k-steps of tile after tile, the ring position carries over
      uint32_t stage = 0, phase = 0;
      for (int tile = t_begin; tile < t_end; ++tile) {
        int t = tile;
        const int tw_i = t % p.tw; t /= p.tw;
        const int th_i = t % p.th; t /= p.th;
        const int td_i = t % p.td;
        const int n_img = t / p.td;
        const int a_d0 = td_i * p.bd * p.istr_d, a_h0 = th_i * p.bh * p.istr_h, a_w0 = tw_i * p.bw * p.istr_w;
        int tap = tap_begin, src = 0, chunk = 0;
        for (int it = 0; it < nstage_iters; ++it) {
          const int g = min(p.group, ksteps - it * p.group);
          mbar_wait(bar_empty + 8u * stage, phase ^ 1u);
          const uint32_t full = bar_full + 8u * stage;
          mbar_expect_tx(full, p.tx_per_kstep * (uint32_t)g);
          for (int j = 0; j < g; ++j) {
            const uint32_t slot = tiles + (stage * p.group + j) * p.slot_bytes;
            const int c_in = chunk * p.ck;
            tma_load_5d(slot, &p.tmA[src], full, c_in, a_w0 + p.tap_ow[tap], a_h0 + p.tap_oh[tap],
                        a_d0 + p.tap_od[tap], n_img);
            tma_load_3d(slot + p.a_alloc, &p.tmB, full, p.src_koff[src] + c_in, n0, (int)p.tap_w[tap]);
            if (++chunk == p.src_chunks[src]) {
              chunk = 0;
              if (++src == p.nsrc) { src = 0; ++tap; }
            }
          }
          if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (elect_one()) {
      // ===== MMA issuer: accumulator buffer = tile parity
      uint32_t stage = 0, phase = 0;
      const uint64_t hi = (uint64_t)p.desc_hi << 32;
      const int k16s = p.ck / 16;
      int ti = 0;
      for (int tile = t_begin; tile < t_end; ++tile, ++ti) {
        const uint32_t buf = (uint32_t)ti & 1u;
        if (ti >= 2) {                                   // the epilogue of tile ti-2 has drained this buffer
          mbar_wait(bar_free + 8u * buf, (uint32_t)((ti >> 1) - 1) & 1u);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        const uint32_t d_tmem = tmem_base + buf * acc_cols;
        uint32_t acc = 0;
        for (int it = 0; it < nstage_iters; ++it) {
          const int g = min(p.group, ksteps - it * p.group);
          mbar_wait(bar_full + 8u * stage, phase);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          for (int j = 0; j < g; ++j) {
            const uint32_t slot = (tiles + (stage * p.group + j) * p.slot_bytes) >> 4;
            const uint32_t slot_b = slot + (p.a_alloc >> 4);
            for (int k = 0; k < k16s; ++k) {
              umma_f16(d_tmem, hi | (uint64_t)(slot + 2u * k), hi | (uint64_t)(slot_b + 2u * k), p.idesc, acc);
              acc = 1;
            }
          }
          umma_commit(bar_empty + 8u * stage);
          if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1u; }
        }
        umma_commit(bar_accum + 8u * buf);
      }
    }
    __syncwarp();
  } else {
    // ===== epilogue warps 2..9
    const int q = warp - 2;
    const int quarter = warp & 3;                        // the TMEM lane quarter this warp may read
    const int r = quarter * 32 + (threadIdx.x & 31);
    const int lw = r % p.bw;
    const int lh = (r / p.bw) % p.bh;
    const int ld = r / (p.bw * p.bh);
    int cb, ce;
    epi_cols(q < 4 ? 0 : 4, p.n_tile, &cb, &ce);
    int ti = 0;
    for (int tile = t_begin; tile < t_end; ++tile, ++ti) {
      const uint32_t buf = (uint32_t)ti & 1u;
      int t = tile;
      const int tw_i = t % p.tw; t /= p.tw;
      const int th_i = t % p.th; t /= p.th;
      const int td_i = t % p.td;
      const int n_img = t / p.td;
      const int d = (td_i * p.bd + ld) * p.ostr_d + p.phase_d[ph], h = (th_i * p.bh + lh) * p.ostr_h + p.phase_h[ph],
                w = (tw_i * p.bw + lw) * p.ostr_w + p.phase_w[ph];
      const bool valid = (ld < p.bd) && d < p.Do && h < p.Ho && w < p.Wo;
      const int64_t vox = (((int64_t)n_img * p.Do + d) * p.Ho + h) * p.Wo + w;
      mbar_wait(bar_accum + 8u * buf, (uint32_t)(ti >> 1) & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t lane_addr = tmem_base + buf * acc_cols + ((uint32_t)(quarter * 32) << 16);
      epilogue_row(p.epi, lane_addr, n0, cb, ce, valid, vox, ksteps == 0);
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (elect_one()) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_free + 8u * buf) : "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2u * acc_cols)
                 : "memory");
  }
}

// fp32 master weights -> bf16 [tap][n_total][k_total]; one weight tensor per produced tensor
struct PackArgs {
  const float* w[M1_MAX_OUT * M1_MAX_SRC];
  int64_t st[M1_MAX_OUT], sr[M1_MAX_OUT], so[M1_MAX_OUT];
  int out_start[M1_MAX_OUT + 1];
  int src_start[M1_MAX_SRC + 1];
  int nout, nsrc, w_by_src;
};
// value of element i of the [tap][n_total][k_total] pack
__device__ __forceinline__ float pack_value(const PackArgs& a, int n_real, int n_total, int k_total, int64_t i) {
  int r = (int)(i % k_total);
  int n = (int)((i / k_total) % n_total);
  const int tap = (int)(i / ((int64_t)k_total * n_total));
  if (n >= n_real) return 0.f;
  int j = 0;
  while (j + 1 < a.nout && n >= a.out_start[j + 1]) ++j;
  n -= a.out_start[j];
  if (a.w_by_src) {
    int s = 0;
    while (s + 1 < a.nsrc && r >= a.src_start[s + 1]) ++s;
    r -= a.src_start[s];
    return a.w[j * a.nsrc + s][tap * a.st[s] + r * a.sr[s] + n * a.so[s]];
  }
  return a.w[j][tap * a.st[j] + r * a.sr[j] + n * a.so[j]];
}

template <typename TW>
__global__ void pack_weights_kernel(const __grid_constant__ PackArgs a, int n_real, int n_total, int k_total,
                                    int taps, TW* __restrict__ out) {
  const int64_t total = (int64_t)taps * n_total * k_total;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x)
    st_f<TW>(out + i, pack_value(a, n_real, n_total, k_total, i));
}

// ---- all packs of a model in ONE launch (the re-pack after every optimizer step: ~350 packs, ~128 M elements)
// A block re-packs one (tap, 32 produced channels, 64 K positions) tile through shared memory: the fp32 master
// weights are read along whichever of their two inner dimensions is contiguous (Conv3D kernels: produced channel
// fastest -> a transpose; Conv3DTranspose kernels and the data-gradient packs: reduction index fastest), the 16-bit
// pack is written K-contiguous, 128 bytes per warp row. All index arithmetic per tile, none per element.
struct PackJob {
  PackArgs a;
  int n_real, n_total, k_total, f16;
  int taps;
  void* out;
};
constexpr int kPackK = 64, kPackN = 32;

__device__ __forceinline__ float pack_value_at(const PackArgs& a, int tap, int r, int n) {
  int j = 0;
  while (j + 1 < a.nout && n >= a.out_start[j + 1]) ++j;
  n -= a.out_start[j];
  if (a.w_by_src) {
    int s = 0;
    while (s + 1 < a.nsrc && r >= a.src_start[s + 1]) ++s;
    r -= a.src_start[s];
    return __ldg(a.w[j * a.nsrc + s] + (tap * a.st[s] + r * a.sr[s] + n * a.so[s]));
  }
  return __ldg(a.w[j] + (tap * a.st[j] + r * a.sr[j] + n * a.so[j]));
}

__global__ void __launch_bounds__(256) pack_tiled_kernel(const PackJob* __restrict__ jobs,
                                                         const int* __restrict__ block_job,
                                                         const int* __restrict__ block_tile) {
  __shared__ float tile[kPackN][kPackK + 1];
  const PackJob& J = jobs[block_job[blockIdx.x]];
  int t = block_tile[blockIdx.x];
  const int kt = (J.k_total + kPackK - 1) / kPackK, nt = (J.n_total + kPackN - 1) / kPackN;
  const int k0 = (t % kt) * kPackK;
  t /= kt;
  const int n0 = (t % nt) * kPackN;
  const int tap = t / nt;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  if (J.a.so[0] == 1) {          // produced channel contiguous in the master weights: a warp reads 32 n of one k
    for (int kk = ty; kk < kPackK; kk += 8) {
      const int k = k0 + kk, n = n0 + tx;
      tile[tx][kk] = (k < J.k_total && n < J.n_real) ? pack_value_at(J.a, tap, k, n) : 0.f;
    }
  } else {                       // reduction index contiguous: a warp reads 2 x 32 k of one n
    for (int nn = ty; nn < kPackN; nn += 8) {
      const int n = n0 + nn;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int k = k0 + tx + 32 * h;
        tile[nn][tx + 32 * h] = (k < J.k_total && n < J.n_real) ? pack_value_at(J.a, tap, k, n) : 0.f;
      }
    }
  }
  __syncthreads();
  for (int nn = ty; nn < kPackN; nn += 8) {
    const int n = n0 + nn, k = k0 + 2 * tx;
    if (n >= J.n_total || k >= J.k_total) continue;         // k_total is a multiple of 16: pairs never straddle
    const int64_t i = ((int64_t)tap * J.n_total + n) * J.k_total + k;
    const float v0 = tile[nn][2 * tx], v1 = tile[nn][2 * tx + 1];
    const uint32_t pk = J.f16 ? pack2<__half>(v0, v1) : pack2<__nv_bfloat16>(v0, v1);
    *reinterpret_cast<uint32_t*>(reinterpret_cast<uint16_t*>(J.out) + i) = pk;
  }
}

struct Plan {
  int ck, n_real, n_total, k_total, n_tile, n_tiles;   // n_total = n_real padded to 16
  int md, mh, mw;                                      // tile grid (produced grid, or one phase of it)
  int sd, sh, sw;                                      // strides
  int nphase;
  int bd, bh, bw, td, th, tw;
  int group, stages;
  uint32_t a_alloc, b_alloc, slot_bytes, smem_bytes, tmem_cols;
};

bool make_plan(const m1_conv_desc* d, Plan* pl) {
  if (!m1_is16(d->act_dtype) || !m1_is16(d->out_dtype)) return false;
  // one operand format per MMA: kind::f16 with different A / B formats is an illegal instruction on sm_100
  if (d->w_dtype != 0 && d->w_dtype != d->act_dtype) return false;
  if (d->nsrc < 1 || d->nsrc > M1_MAX_SRC || d->nout < 1 || d->nout > M1_MAX_OUT) return false;
  const int taps = d->kernel[0] * d->kernel[1] * d->kernel[2];
  if (taps > 32) return false;
  int nphase = 1;
  for (int i = 0; i < 3; ++i) {
    if (d->stride[i] < 1 || d->stride[i] > 2) return false;
    if (d->mode == M1_CONV_TRANSPOSED) nphase *= d->stride[i];
  }
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
  // N tile: largest divisor of n_total that is a multiple of 16 and <= 256
  int n_tile = 0;
  for (int c = 256; c >= 16; c -= 16)
    if (n_total % c == 0) { n_tile = c; break; }
  if (!n_tile) return false;
  // tile grid: the produced grid (FWD) or one output phase of it (TRANSPOSED)
  int M[3];
  for (int i = 0; i < 3; ++i)
    M[i] = d->mode == M1_CONV_TRANSPOSED ? (d->out_dhw[i] + d->stride[i] - 1) / d->stride[i] : d->out_dhw[i];
  const int D = M[0], H = M[1], W = M[2];
  // TMA box extent over the gathered grid: brick * element stride <= 256 per dimension
  const int is_d = d->mode == M1_CONV_FWD ? d->stride[0] : 1, is_h = d->mode == M1_CONV_FWD ? d->stride[1] : 1,
            is_w = d->mode == M1_CONV_FWD ? d->stride[2] : 1;
  // brick: maximise useful voxels / 128 over all (bd,bh,bw) with bd*bh*bw <= 128
  double best = -1;
  int bbd = 1, bbh = 1, bbw = 1;
  for (int bd = 1; bd <= 128 && bd <= D; ++bd)
    for (int bh = 1; bd * bh <= 128 && bh <= H; ++bh) {
      int bw = 128 / (bd * bh);
      if (bw > W) bw = W;
      if (bw < 1) continue;
      if (bd * is_d > 256 || bh * is_h > 256 || bw * is_w > 256) continue;
      const int64_t tiles = (int64_t)((D + bd - 1) / bd) * ((H + bh - 1) / bh) * ((W + bw - 1) / bw);
      const double eff = (double)D * H * W / (128.0 * tiles);
      // prefer wide bricks (longer contiguous runs) on ties
      const double score = eff + 1e-6 * bw;
      if (score > best) { best = score; bbd = bd; bbh = bh; bbw = bw; }
    }
  if (best < 0) return false;
  pl->ck = ck; pl->n_real = n_real; pl->n_total = n_total; pl->k_total = k_total; pl->n_tile = n_tile;
  pl->n_tiles = n_total / n_tile;
  pl->md = D; pl->mh = H; pl->mw = W;
  pl->sd = d->stride[0]; pl->sh = d->stride[1]; pl->sw = d->stride[2];
  pl->nphase = nphase;
  pl->bd = bbd; pl->bh = bbh; pl->bw = bbw;
  pl->td = (D + bbd - 1) / bbd; pl->th = (H + bbh - 1) / bbh; pl->tw = (W + bbw - 1) / bbw;
  pl->a_alloc = 128u * ck * 2u;
  pl->b_alloc = ((uint32_t)n_tile * ck * 2u + 1023u) & ~1023u;
  pl->slot_bytes = pl->a_alloc + pl->b_alloc;
  uint32_t cols = 32;
  while ((int)cols < n_tile) cols <<= 1;
  pl->tmem_cols = cols;
  // k-steps per stage: aim at >= 64 channels of work per barrier round trip
  pl->group = 64 / ck;
  const int ksteps = ((taps + nphase - 1) / nphase) * (k_total / ck);   // typical k-steps of one CTA
  // co-residency target: as many CTAs/SM as TMEM allows (<=4), smem split accordingly
  int ctas = 512 / (int)cols;
  if (ctas > 4) ctas = 4;
  const uint32_t budget = (227u * 1024u) / ctas - 2048u;
  int stages = (int)(budget / (pl->slot_bytes * pl->group));
  if (stages > 12) stages = 12;
  const int iters = std::max(1, (ksteps + pl->group - 1) / pl->group);
  if (stages > iters) stages = iters;
  if (stages < 2) {
    // not enough room at this co-residency: fall back to 1 CTA/SM
    stages = (int)((227u * 1024u - 2048u) / (pl->slot_bytes * pl->group));
    if (stages > 12) stages = 12;
    if (stages > iters) stages = std::max(iters, 1);
    if (stages < 1) return false;
  }
  pl->stages = stages;
  pl->smem_bytes = 2048u + (uint32_t)stages * pl->group * pl->slot_bytes;
  return true;
}

}  // namespace

extern "C" int m1_conv3d_tc_supported(const m1_conv_desc* d) {
  Plan pl;
  return make_plan(d, &pl) ? 1 : 0;
}

extern "C" int64_t m1_conv3d_packed_bytes(const m1_conv_desc* d) {
  Plan pl;
  if (!make_plan(d, &pl)) return 0;
  const int taps = d->kernel[0] * d->kernel[1] * d->kernel[2];
  return (int64_t)taps * pl.n_total * pl.k_total * 2;
}

static void fill_pack_args(const m1_conv_desc* d, const float* const* w, PackArgs* a) {
  memset(a, 0, sizeof(*a));
  a->nout = d->nout;
  a->nsrc = d->nsrc;
  a->w_by_src = d->w_by_src;
  int acc = 0;
  for (int j = 0; j < d->nout; ++j) {
    a->out_start[j] = acc;
    acc += d->out_c[j];
  }
  a->out_start[d->nout] = acc;
  acc = 0;
  for (int s = 0; s < d->nsrc; ++s) {
    a->src_start[s] = acc;
    acc += d->src_c[s];
  }
  a->src_start[d->nsrc] = acc;
  for (int j = 0; j < M1_MAX_OUT; ++j) {
    a->st[j] = d->w_stride_tap[j]; a->sr[j] = d->w_stride_red[j]; a->so[j] = d->w_stride_out[j];
  }
  const int nw = d->w_by_src ? d->nout * d->nsrc : d->nout;
  for (int i = 0; i < nw; ++i) a->w[i] = w[i];
}

extern "C" int m1_conv3d_pack_weights(m1_ctx* ctx, const m1_conv_desc* d, const float* const* w,
                                      void* w_packed, void* stream) {
  Plan pl;
  M1_CHECK(make_plan(d, &pl), "m1_conv3d_pack_weights: launch not supported by the tcgen05 engine");
  const int taps = d->kernel[0] * d->kernel[1] * d->kernel[2];
  const int64_t total = (int64_t)taps * pl.n_total * pl.k_total;
  const int blocks = (int)std::min<int64_t>(cdiv64(total, 256), 148 * 8);
  PackArgs a;
  fill_pack_args(d, w, &a);
  if (m1_conv_w_dtype(d) == M1_F16)
    pack_weights_kernel<__half><<<blocks, 256, 0, (cudaStream_t)stream>>>(a, pl.n_real, pl.n_total, pl.k_total, taps,
                                                                        reinterpret_cast<__half*>(w_packed));
  else
    pack_weights_kernel<__nv_bfloat16><<<blocks, 256, 0, (cudaStream_t)stream>>>(
        a, pl.n_real, pl.n_total, pl.k_total, taps, reinterpret_cast<__nv_bfloat16*>(w_packed));
  M1_LAUNCH_CHECK(ctx);
  return 0;
}

struct m1_pack_plan {
  PackJob* jobs;
  int* block_job;
  int* block_tile;
  int nblocks;
};

extern "C" int m1_pack_plan_create(m1_ctx* ctx, int njobs, const m1_conv_desc* const* descs,
                                   const float* const* const* ws, void* const* packed, m1_pack_plan** out) {
  M1_CHECK(ctx && out && njobs > 0, "m1_pack_plan_create: bad arguments");
  std::vector<PackJob> jobs((size_t)njobs);
  std::vector<int> bj, bt;
  for (int q = 0; q < njobs; ++q) {
    Plan pl;
    M1_CHECK(make_plan(descs[q], &pl), "m1_pack_plan_create: job %d is not a tcgen05 launch", q);
    PackJob& J = jobs[(size_t)q];
    fill_pack_args(descs[q], ws[q], &J.a);
    J.taps = descs[q]->kernel[0] * descs[q]->kernel[1] * descs[q]->kernel[2];
    J.n_real = pl.n_real; J.n_total = pl.n_total; J.k_total = pl.k_total;
    J.f16 = m1_conv_w_dtype(descs[q]) == M1_F16;
    J.out = packed[q];
    M1_CHECK(((uintptr_t)packed[q] & 3) == 0, "m1_pack_plan_create: pack %d not 4-byte aligned", q);
    const int tiles = J.taps * ((pl.n_total + kPackN - 1) / kPackN) * ((pl.k_total + kPackK - 1) / kPackK);
    for (int t = 0; t < tiles; ++t) {
      bj.push_back(q);
      bt.push_back(t);
    }
  }
  m1_pack_plan* p = new m1_pack_plan();
  p->nblocks = (int)bj.size();
  M1_CUDA(cudaMalloc(&p->jobs, jobs.size() * sizeof(PackJob)));
  M1_CUDA(cudaMalloc(&p->block_job, bj.size() * sizeof(int)));
  M1_CUDA(cudaMalloc(&p->block_tile, bt.size() * sizeof(int)));
  M1_CUDA(cudaMemcpy(p->jobs, jobs.data(), jobs.size() * sizeof(PackJob), cudaMemcpyHostToDevice));
  M1_CUDA(cudaMemcpy(p->block_job, bj.data(), bj.size() * sizeof(int), cudaMemcpyHostToDevice));
  M1_CUDA(cudaMemcpy(p->block_tile, bt.data(), bt.size() * sizeof(int), cudaMemcpyHostToDevice));
  *out = p;
  return 0;
}

extern "C" int m1_pack_plan_run(m1_ctx* ctx, const m1_pack_plan* plan, void* stream) {
  M1_CHECK(ctx && plan, "m1_pack_plan_run: NULL argument");
  pack_tiled_kernel<<<(unsigned)plan->nblocks, 256, 0, (cudaStream_t)stream>>>(plan->jobs, plan->block_job,
                                                                                plan->block_tile);
  M1_LAUNCH_CHECK(ctx);
  return 0;
}

extern "C" int m1_pack_plan_destroy(m1_pack_plan* plan) {
  if (!plan) return 0;
  cudaFree(plan->jobs);
  cudaFree(plan->block_job);
  cudaFree(plan->block_tile);
  delete plan;
  return 0;
}

// which = 0: per-tap engine. out: ck, n_tile, n_tiles, bd, bh, bw, stages, group, smem_bytes, tmem_cols, ctas
int m1_conv3d_tc_plan_info(const m1_conv_desc* d, int32_t* out) {
  Plan pl;
  if (!make_plan(d, &pl)) return 0;
  const int32_t v[11] = {pl.ck, pl.n_tile, pl.n_tiles, pl.bd, pl.bh, pl.bw, pl.stages, pl.group,
                         (int32_t)pl.smem_bytes, (int32_t)pl.tmem_cols,
                         (int32_t)((int64_t)d->batch * pl.td * pl.th * pl.tw * pl.n_tiles * pl.nphase)};
  for (int i = 0; i < 11; ++i) out[i] = v[i];
  return 11;
}

extern "C" int m1_conv3d_halo_engine(const m1_conv_desc* d) {
  int pref = 0;
  if (d->tune[0] == 1 || d->tune[0] == 3 || !m1_conv3d_halo_supported(d, &pref)) return 0;
  return (d->tune[0] == 2 || pref) ? 1 : 0;
}

int m1_conv3d_tc(m1_ctx* ctx, const m1_conv_desc* d, const void* const* srcs,
                 const void* w_packed, const float* const* bias, void* const* outs,
                 cudaStream_t st) {
  if (m1_conv3d_halo_engine(d)) return m1_conv3d_halo(ctx, d, srcs, w_packed, bias, outs, st);
  Plan pl;
  M1_CHECK(make_plan(d, &pl), "m1_conv3d: launch not supported by the tcgen05 engine");
  M1_CHECK(w_packed != nullptr, "m1_conv3d: tcgen05 engine needs the bf16 weight pack");
  M1_CHECK(ctx->encode_tiled != nullptr, "cuTensorMapEncodeTiled entry point unavailable");
  EncodeTiledFn encode = reinterpret_cast<EncodeTiledFn>(ctx->encode_tiled);

  static_assert(sizeof(TcParams) < 4000, "kernel parameter block too large");
  TcParams p;
  memset(&p, 0, sizeof(p));
  const CUtensorMapSwizzle swz = swizzle_for(pl.ck);
  const bool fwd = d->mode == M1_CONV_FWD;
  const int is[3] = {fwd ? pl.sd : 1, fwd ? pl.sh : 1, fwd ? pl.sw : 1};   // gathered-grid step per tile voxel
  const int os[3] = {fwd ? 1 : pl.sd, fwd ? 1 : pl.sh, fwd ? 1 : pl.sw};   // produced-grid step per tile voxel
  const int Di = d->in_dhw[0], Hi = d->in_dhw[1], Wi = d->in_dhw[2];
  int koff = 0;
  for (int s = 0; s < d->nsrc; ++s) {
    const cuuint64_t C = (cuuint64_t)d->src_c[s];
    cuuint64_t dims[5] = {C, (cuuint64_t)Wi, (cuuint64_t)Hi, (cuuint64_t)Di, (cuuint64_t)d->batch};
    cuuint64_t strides[4] = {C * 2, C * 2 * Wi, C * 2 * Wi * Hi, C * 2 * Wi * Hi * Di};
    // strided gather: box extent b*s with element stride s loads b voxels (every s-th)
    cuuint32_t box[5] = {(cuuint32_t)pl.ck, (cuuint32_t)(pl.bw * is[2]), (cuuint32_t)(pl.bh * is[1]),
                         (cuuint32_t)(pl.bd * is[0]), 1};
    cuuint32_t es[5] = {1, (cuuint32_t)is[2], (cuuint32_t)is[1], (cuuint32_t)is[0], 1};
    M1_CHECK(((uintptr_t)srcs[s] & 15) == 0, "m1_conv3d: gathered tensor %d not 16-byte aligned", s);
    CUresult r = encode(&p.tmA[s], tm_dtype(d->act_dtype), 5, const_cast<void*>(srcs[s]),
                        dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    M1_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(A, src %d) failed: %d", s, (int)r);
    p.src_chunks[s] = d->src_c[s] / pl.ck;
    p.src_koff[s] = koff;
    koff += d->src_c[s];
  }
  const int taps = d->kernel[0] * d->kernel[1] * d->kernel[2];
  {
    cuuint64_t dims[3] = {(cuuint64_t)pl.k_total, (cuuint64_t)pl.n_total, (cuuint64_t)taps};
    cuuint64_t strides[2] = {(cuuint64_t)pl.k_total * 2, (cuuint64_t)pl.k_total * 2 * pl.n_total};
    cuuint32_t box[3] = {(cuuint32_t)pl.ck, (cuuint32_t)pl.n_tile, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = encode(&p.tmB, tm_dtype(m1_conv_w_dtype(d)), 3, const_cast<void*>(w_packed),
                        dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    M1_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(B) failed: %d", (int)r);
  }
  p.nsrc = d->nsrc;
  // ---- tap table (and output phases of a transposed gather)
  {
    const int K[3] = {d->kernel[0], d->kernel[1], d->kernel[2]};
    const int S[3] = {pl.sd, pl.sh, pl.sw};
    const int P[3] = {d->pad[0], d->pad[1], d->pad[2]};
    int n = 0, ph = 0;
    const int pd_n = fwd ? 1 : S[0], ph_n = fwd ? 1 : S[1], pw_n = fwd ? 1 : S[2];
    for (int fd = 0; fd < pd_n; ++fd)
      for (int fh = 0; fh < ph_n; ++fh)
        for (int fw = 0; fw < pw_n; ++fw) {
          p.phase_tap0[ph] = (uint8_t)n;
          p.phase_d[ph] = (int8_t)fd; p.phase_h[ph] = (int8_t)fh; p.phase_w[ph] = (int8_t)fw;
          const int F[3] = {fd, fh, fw};
          for (int a = 0; a < K[0]; ++a)
            for (int b = 0; b < K[1]; ++b)
              for (int c = 0; c < K[2]; ++c) {
                const int kk[3] = {a, b, c};
                int off[3];
                bool ok = true;
                for (int i = 0; i < 3; ++i) {
                  if (fwd) {
                    off[i] = kk[i] - P[i];
                  } else {
                    const int num = F[i] + P[i] - kk[i];
                    if (((num % S[i]) + S[i]) % S[i] != 0) { ok = false; break; }
                    off[i] = num / S[i];     // exact
                  }
                }
                if (!ok) continue;
                p.tap_w[n] = (uint8_t)((a * K[1] + b) * K[2] + c);
                p.tap_od[n] = (int8_t)off[0]; p.tap_oh[n] = (int8_t)off[1]; p.tap_ow[n] = (int8_t)off[2];
                ++n;
              }
          ++ph;
        }
    p.phase_tap0[ph] = (uint8_t)n;
    p.nphase = ph;
  }
  p.istr_d = is[0]; p.istr_h = is[1]; p.istr_w = is[2];
  p.ostr_d = os[0]; p.ostr_h = os[1]; p.ostr_w = os[2];
  p.bd = pl.bd; p.bh = pl.bh; p.bw = pl.bw;
  p.td = pl.td; p.th = pl.th; p.tw = pl.tw;
  p.Do = d->out_dhw[0]; p.Ho = d->out_dhw[1]; p.Wo = d->out_dhw[2];
  p.n_tile = pl.n_tile;
  p.ck = pl.ck;
  p.group = pl.group;
  p.stages = pl.stages;
  p.a_alloc = pl.a_alloc;
  p.slot_bytes = pl.slot_bytes;
  p.tx_per_kstep = (uint32_t)(pl.bd * pl.bh * pl.bw) * pl.ck * 2u + (uint32_t)pl.n_tile * pl.ck * 2u;
  p.tmem_cols = pl.tmem_cols;
  // instruction descriptor: D=f32, A / B format (f16 or bf16, independently), both K-major, N>>3 at [17,23),
  // M>>4 at [24,29)
  p.idesc = (1u << 4) | (idesc_fmt(d->act_dtype) << 7) | (idesc_fmt(m1_conv_w_dtype(d)) << 10) |
            ((uint32_t)(pl.n_tile >> 3) << 17) | ((128u >> 4) << 24);
  // smem descriptor high word: SBO = 8 rows * ck*2 bytes (>>4) at [32,46), version 1 at [46,48),
  // layout type at [61,64): 2 = SW128, 4 = SW64, 6 = SW32; LBO (unused for swizzled K-major) = 1
  const uint32_t sbo = (8u * pl.ck * 2u) >> 4;
  const uint32_t layout = pl.ck == 64 ? 2u : pl.ck == 32 ? 4u : 6u;
  p.desc_hi = sbo | (1u << 14) | (layout << 29);
  M1_CHECK(epi_fill(&p.epi, d, bias, outs), "m1_conv3d: too many produced channels for the tcgen05 epilogue table");
  for (int j = 0; j < d->nout; ++j)
    M1_CHECK(((uintptr_t)outs[j] & 15) == 0, "m1_conv3d: produced tensor %d not 16-byte aligned", j);

  const uint32_t smem_bytes = pl.smem_bytes;
  static const int g_multi = getenv("M1_CONV_MULTI") ? atoi(getenv("M1_CONV_MULTI")) : 0;
  if (d->tune[0] == 3 || g_multi) {
    // multi-tile variant (see conv_tc_multi_kernel): the producer streams k-steps ACROSS tile boundaries, so the
    // depth of the ring is chosen for memory latency (>= ~50 KB in flight per SM), not for the k-steps of one
    // tile - with a ring as deep as one tile (1 stage for a 1x1x1 convolution) every tile would pay the full
    // TMA round trip again, which is exactly what bounds the single-tile kernel on short-K launches.
    uint32_t acc_cols = 32;
    while ((int)acc_cols < pl.n_tile) acc_cols <<= 1;
    const int ksteps_tile = std::max(1, ((taps + pl.nphase - 1) / pl.nphase) * (pl.k_total / pl.ck));
    const int mgroup = std::min(pl.group, ksteps_tile);
    int ctas_sm = std::min(3, std::max(1, 512 / (int)(2u * acc_cols)));          // 320 threads x 64 registers: <= 3
    int mstages = (int)(((227u * 1024u) / ctas_sm - 2048u) / (pl.slot_bytes * mgroup));
    mstages = std::min(mstages, 8);
    M1_CHECK(mstages >= 2, "m1_conv3d: multi-tile variant needs >= 2 pipeline stages (slot %u bytes x %d)",
             pl.slot_bytes, mgroup);
    p.group = mgroup;
    p.stages = mstages;
    const uint32_t msmem = 2048u + (uint32_t)mstages * mgroup * pl.slot_bytes;
    const int total_tiles = d->batch * pl.td * pl.th * pl.tw;
    const int64_t ctas1 = (int64_t)total_tiles * pl.n_tiles * pl.nphase;
    // tiles per CTA: a few waves of CTAs over the machine, each long enough to amortise its set-up
    int per = (int)std::max<int64_t>(1, std::min<int64_t>(32, ctas1 / ((int64_t)ctx->num_sms * ctas_sm * 2)));
    static int multi_set = 0;
    if (!multi_set) {
      M1_CUDA(cudaFuncSetAttribute(conv_tc_multi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      multi_set = 1;
    }
    dim3 mgrid((unsigned)((total_tiles + per - 1) / per), (unsigned)pl.n_tiles, (unsigned)pl.nphase);
    conv_tc_multi_kernel<<<mgrid, kThreadsMulti, msmem, st>>>(p, per, total_tiles, acc_cols);
    M1_LAUNCH_CHECK(ctx);
    return 0;
  }
  static int smem_set = 0;
  if (smem_set < (int)pl.smem_bytes) {
    M1_CUDA(cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 227 * 1024));
    smem_set = 227 * 1024;
  }
  dim3 grid((unsigned)(d->batch * pl.td * pl.th * pl.tw), (unsigned)pl.n_tiles, (unsigned)pl.nphase);
  conv_tc_kernel<<<grid, kThreads, smem_bytes, st>>>(p);
  M1_LAUNCH_CHECK(ctx);
  return 0;
}
