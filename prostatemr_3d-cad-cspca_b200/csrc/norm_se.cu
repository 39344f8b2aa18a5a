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
#include "tc_common.cuh"
#include <algorithm>
#include <cstdlib>
#include <unordered_set>

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
  // the same vector from SHARED memory (32-bit shared-window address): the staged skeletons below
  __device__ __forceinline__ void load_shared(uint32_t saddr) {
    if constexpr (ES * VW == 32) {
      asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "r"(saddr));
      asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]) : "r"(saddr + 16u));
    } else if constexpr (ES * VW == 16) {
      asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "r"(saddr));
    } else if constexpr (ES * VW == 8) {
      asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(w[0]), "=r"(w[1]) : "r"(saddr));
    } else if constexpr (ES == 4) {
      asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w[0]) : "r"(saddr));
    } else {
      uint16_t h;
      asm volatile("ld.shared.b16 %0, [%1];" : "=h"(h) : "r"(saddr));
      w[0] = h;
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

// PIPE: software-pipelined main loop - the raw loads of iteration i+1 are issued BEFORE iteration i is computed
// (two register sets, ping-pong). The gate kernels spend ~500 instructions per iteration between load batches; ncu
// shows them latency-bound (long-scoreboard stalls, issue slots 44 % busy at 25 % occupancy).
template <int UNR, int NT, int ES, int VW, typename G>
__device__ __forceinline__ void pipelined_rows(const void* const (&src)[NT], int64_t& r, int64_t& off, int64_t r1,
                                               int lanes, int64_t step, G g) {
  RawVec<ES, VW> A[UNR][NT], B[UNR][NT];
  auto load = [&](RawVec<ES, VW> (&dst)[UNR][NT], int64_t o) {
#pragma unroll
    for (int u = 0; u < UNR; ++u)
#pragma unroll
      for (int t = 0; t < NT; ++t) dst[u][t].load(src[t], o + u * step);
  };
  auto full = [&](int64_t rr) { return rr + (int64_t)(UNR - 1) * lanes < r1; };
  const int64_t dr = (int64_t)UNR * lanes, doff = UNR * step;
  if (!full(r)) return;
  load(A, off);
  while (true) {
    const bool nb = full(r + dr);
    if (nb) load(B, off + doff);
#pragma unroll
    for (int u = 0; u < UNR; ++u) g(r + (int64_t)u * lanes, off + u * step, A[u]);
    r += dr; off += doff;
    if (!nb) break;
    const bool na = full(r + dr);
    if (na) load(A, off + doff);
#pragma unroll
    for (int u = 0; u < UNR; ++u) g(r + (int64_t)u * lanes, off + u * step, B[u]);
    r += dr; off += doff;
    if (!na) break;
  }
}

template <int K, int VW, int UNR, int NT, int ES, bool PIPE = false, typename P, typename F, typename FIN>
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
      if constexpr (PIPE) {
        pipelined_rows<UNR, NT, ES, VW>(src, r, off, r1, lanes, step,
                                        [&](int64_t row, int64_t, const RawVec<ES, VW> (&raw)[NT]) {
                                          f(row, cbase, regs, raw, acc);
                                        });
      } else {
        for (; r + (int64_t)(UNR - 1) * lanes < r1; r += (int64_t)UNR * lanes, off += UNR * step) {
          RawVec<ES, VW> raw[UNR][NT];
#pragma unroll
          for (int u = 0; u < UNR; ++u)
#pragma unroll
            for (int t = 0; t < NT; ++t) raw[u][t].load(src[t], off + u * step);
#pragma unroll
          for (int u = 0; u < UNR; ++u) f(r + (int64_t)u * lanes, cbase, regs, raw[u], acc);
        }
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
template <int VW, int UNR, int NT, int ES, bool PIPE = false, typename P, typename F>
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
    if constexpr (PIPE) {
      pipelined_rows<UNR, NT, ES, VW>(src, r, off, r1, lanes, step,
                                      [&](int64_t row, int64_t o, const RawVec<ES, VW> (&raw)[NT]) {
                                        f(row, o, cbase, regs, raw);
                                      });
    } else {
      for (; r + (int64_t)(UNR - 1) * lanes < r1; r += (int64_t)UNR * lanes, off += UNR * step) {
        RawVec<ES, VW> raw[UNR][NT];
#pragma unroll
        for (int u = 0; u < UNR; ++u)
#pragma unroll
          for (int t = 0; t < NT; ++t) raw[u][t].load(src[t], off + u * step);
#pragma unroll
        for (int u = 0; u < UNR; ++u) f(r + (int64_t)u * lanes, off + u * step, cbase, regs, raw[u]);
      }
    }
    for (; r < r1; r += lanes, off += step) {
      RawVec<ES, VW> raw[NT];
#pragma unroll
      for (int t = 0; t < NT; ++t) raw[t].load(src[t], off);
      f(r, off, cbase, regs, raw);
    }
  }
}

// ---- STAGED skeletons: the same two traversals fed through a bulk-copy ring in shared memory ----------------
// The register skeletons above keep UNR x NT vector loads in flight per thread; with the 90-130 registers the
// backward kernels need that is ~50 KB per SM, and ncu shows them at 1.7-2.9 TB/s (long-scoreboard bound). Here one
// elected thread of warp 0 streams the slab of every tensor into a ring of `stages` x NT x 8 KB with
// cp.async.bulk (a slab of rows of an NDHWC tensor is ONE contiguous byte range), completion on an mbarrier per
// stage, always `stages - 1` chunks ahead of the consumers; all eight warps read their rows from shared memory
// (conflict-free: consecutive threads, consecutive 8 / 16 bytes) and hand the stage back as soon as the rows are in
// registers. Bytes in flight per SM = ring size (2 blocks x 64-96 KB), whatever the register budget.
// Eligibility (host, staged_plan): C * element size a multiple of 16 bytes, C / VW a divisor of 256 - every tensor
// of the tensor-core modes; odd shapes keep the register skeletons.
constexpr int kStageBytes = 8192;        // bytes of ONE tensor in one stage
constexpr int kMaxStages = 8;

// barriers of the ring: full[kMaxStages] | empty[kMaxStages] in the first 128 bytes, tiles after them
__device__ __forceinline__ void ring_init(uint8_t* dsm, int stages, uint32_t* full, uint32_t* empty, uint32_t* data) {
  const uint32_t base = (tc::smem_u32(dsm) + 127u) & ~127u;
  *full = base;
  *empty = base + 8u * kMaxStages;
  *data = base + 128u;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      tc::mbar_init(*full + 8u * s, 1);
      tc::mbar_init(*empty + 8u * s, TB / 32);       // one arrival per warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
}

// Rows [r0, r1) of the NT tensors through the ring; per chunk: the thread's rows -> registers, stage released,
// then G(row, element offset inside the sample, raw[NT]) per row. All TB threads must call.
template <int VW, int NT, int ES, typename G>
__device__ __forceinline__ void ring_run(const void* const (&src)[NT], int64_t r0, int64_t r1, int C, int lanes,
                                         int my_lane, int cbase, int stages, uint32_t full, uint32_t empty,
                                         uint32_t data, G g) {
  constexpr int M = kStageBytes / (TB * VW * ES);      // rows per thread and chunk
  static_assert(M >= 1, "stage too small for this vector width");
  const int rows_chunk = M * lanes;
  const int64_t row_bytes = (int64_t)C * ES;
  bool producer = false;
  if (threadIdx.x < 32) producer = tc::elect_one() != 0;
  int ps = 0;
  uint32_t pph = 0;
  int64_t pr = r0;                                      // first row of the next chunk to request
  auto issue = [&]() {
    const uint32_t bytes = (uint32_t)(min((int64_t)rows_chunk, r1 - pr) * row_bytes);
    tc::mbar_wait(empty + 8u * ps, pph ^ 1u);
    tc::mbar_expect_tx(full + 8u * ps, bytes * NT);
#pragma unroll
    for (int t = 0; t < NT; ++t)
      tc::bulk_load_1d(data + (uint32_t)(ps * NT + t) * kStageBytes, reinterpret_cast<const char*>(src[t]) + pr * row_bytes,
                       bytes, full + 8u * ps);
    pr += rows_chunk;
    if (++ps == stages) { ps = 0; pph ^= 1u; }
  };
  if (producer)
    for (int k = 0; k < stages - 1 && pr < r1; ++k) issue();
  int s = 0;
  uint32_t ph = 0;
  for (int64_t r = r0; r < r1; r += rows_chunk) {
    if (producer && pr < r1) issue();                   // refills the stage every warp released one iteration ago
    const int nrows = (int)min((int64_t)rows_chunk, r1 - r);
    tc::mbar_wait(full + 8u * s, ph);
    const uint32_t st = data + (uint32_t)s * (NT * kStageBytes);
    RawVec<ES, VW> raw[M][NT];
#pragma unroll
    for (int j = 0; j < M; ++j) {
      const int rr = my_lane + j * lanes;
      if (rr < nrows) {
#pragma unroll
        for (int t = 0; t < NT; ++t) raw[j][t].load_shared(st + (uint32_t)t * kStageBytes + (uint32_t)(rr * C + cbase) * ES);
      }
    }
    __syncwarp();
    if ((threadIdx.x & 31) == 0) tc::mbar_arrive(empty + 8u * s);
#pragma unroll
    for (int j = 0; j < M; ++j) {
      const int rr = my_lane + j * lanes;
      if (rr < nrows) g(r + rr, (r + rr) * C + cbase, raw[j]);
    }
    if (++s == stages) { s = 0; ph ^= 1u; }
  }
}

template <int VW, int NT, int ES, typename P, typename F>
__device__ __forceinline__ void stream_rows_staged(const void* const (&src)[NT], int64_t voxels, int C,
                                                   int64_t rows_per_slab, int stages, uint8_t* dsm, P prep, F f) {
  uint32_t full, empty, data;
  ring_init(dsm, stages, &full, &empty, &data);
  const int CG = C / VW, lanes = TB / CG;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_slab;
  const int64_t r1 = min(r0 + rows_per_slab, voxels);
  const int cbase = (threadIdx.x % CG) * VW, my_lane = threadIdx.x / CG;
  const auto regs = prep(cbase);
  ring_run<VW, NT, ES>(src, r0, r1, C, lanes, my_lane, cbase, stages, full, empty, data,
                       [&](int64_t row, int64_t off, const RawVec<ES, VW> (&raw)[NT]) { f(row, off, cbase, regs, raw); });
}

template <int K, int VW, int NT, int ES, typename P, typename F, typename FIN>
__device__ __forceinline__ void reduce_rows_staged(const void* const (&src)[NT], int64_t voxels, int C,
                                                   int64_t rows_per_slab, int stages, uint8_t* dsm, ReduceScratch rs,
                                                   P prep, F f, FIN fin) {
  uint32_t full, empty, data;
  ring_init(dsm, stages, &full, &empty, &data);
  const int CG = C / VW, lanes = TB / CG;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_slab;
  const int64_t r1 = min(r0 + rows_per_slab, voxels);
  // fold geometry of reduce_rows (cgs == CG here; CG divides 256: a power of two)
  const bool fold = CG < 32;
  const int nslots = fold ? TB / 32 : lanes;
  float* smem = reinterpret_cast<float*>(dsm + (data - tc::smem_u32(dsm)));     // the ring, once it has drained
  float acc[K][8];
#pragma unroll
  for (int k = 0; k < K; ++k)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[k][i] = 0.f;
  const int cbase = (threadIdx.x % CG) * VW, my_lane = threadIdx.x / CG;
  {
    const auto regs = prep(cbase);
    ring_run<VW, NT, ES>(src, r0, r1, C, lanes, my_lane, cbase, stages, full, empty, data,
                         [&](int64_t row, int64_t, const RawVec<ES, VW> (&raw)[NT]) { f(row, cbase, regs, raw, acc); });
  }
  if (fold) {
#pragma unroll
    for (int k = 0; k < K; ++k)
#pragma unroll
      for (int i = 0; i < VW; ++i)
        for (int o = CG; o < 32; o <<= 1) acc[k][i] += __shfl_xor_sync(0xffffffffu, acc[k][i], o);
  }
  __syncthreads();               // every stage consumed: the ring memory is free for the per-slot partial sums
  if (!fold || (int)(threadIdx.x & 31) < CG) {
    const int slot = fold ? (int)(threadIdx.x >> 5) : my_lane;
#pragma unroll
    for (int k = 0; k < K; ++k)
#pragma unroll
      for (int i = 0; i < VW; ++i) smem[(slot * K + k) * C + cbase + i] = acc[k][i];
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
template <typename T, int VW, bool ST>
__global__ void __launch_bounds__(TB) inorm_stats_kernel(const T* __restrict__ x, int64_t voxels, int C,
                                                                   int64_t rows_per_slab, float inv_v, float eps,
                                                                   ReduceScratch rs, float* __restrict__ stats,
                                                                   int stages) {
  extern __shared__ __align__(16) uint8_t dsm[];
  const int n = blockIdx.y;
  const void* const src[1] = {x + (int64_t)n * voxels * C};
  float* st = stats + (int64_t)n * C * 2;
  auto prep = [](int) { return 0; };
  auto body = [&](int64_t, int, int, const RawVec<sizeof(T), VW> (&raw)[1], float (&acc)[2][8]) {
    float v[8];
    raw[0].template unpack<T>(v);
#pragma unroll
    for (int i = 0; i < VW; ++i) { acc[0][i] += v[i]; acc[1][i] = fmaf(v[i], v[i], acc[1][i]); }
  };
  auto fin = [&](int c, const float (&t)[2]) {
    const float mean = t[0] * inv_v;
    const float var = fmaxf(t[1] * inv_v - mean * mean, 0.f);
    st[2 * c] = mean;
    st[2 * c + 1] = rsqrtf(var + eps);
  };
  if constexpr (ST)
    reduce_rows_staged<2, VW, 1, sizeof(T)>(src, voxels, C, rows_per_slab, stages, dsm, rs, prep, body, fin);
  else
    reduce_rows<2, VW, 8, 1, sizeof(T)>(src, voxels, C, rows_per_slab, reinterpret_cast<float*>(dsm), rs, prep, body, fin);
}

template <typename T, int VW, bool ST>
__global__ void __launch_bounds__(TB) inorm_act_fwd_kernel(const T* __restrict__ x, const float* __restrict__ stats,
                                                                     const float* __restrict__ gamma,
                                                                     const float* __restrict__ beta, int64_t voxels, int C,
                                                                     float slope, T* __restrict__ y,
                                                                     __nv_bfloat16* __restrict__ y2, int64_t rows_per_slab,
                                                                     int stages) {
  extern __shared__ __align__(16) uint8_t dsm[];
  const int n = blockIdx.y;
  const void* const src[1] = {x + (int64_t)n * voxels * C};
  T* yb = y + (int64_t)n * voxels * C;
  __nv_bfloat16* y2b = y2 ? y2 + (int64_t)n * voxels * C : nullptr;      // optional bf16 twin of the output
  auto prep = [&](int cbase) { return norm_regs<VW>(stats + (int64_t)n * C * 2, gamma, beta, cbase); };
  auto body = [&](int64_t, int64_t off, int, const NormRegs& q, const RawVec<sizeof(T), VW> (&raw)[1]) {
    float v[8];
    raw[0].template unpack<T>(v);
#pragma unroll
    for (int k = 0; k < VW; ++k) v[k] = lrelu(fmaf(v[k], q.a[k], q.b[k]), slope);
    stv<T, VW>(yb + off, v);
    if (y2b) stv<__nv_bfloat16, VW>(y2b + off, v);
  };
  if constexpr (ST) stream_rows_staged<VW, 1, sizeof(T)>(src, voxels, C, rows_per_slab, stages, dsm, prep, body);
  else stream_rows<VW, 8, 1, sizeof(T)>(src, voxels, C, rows_per_slab, prep, body);
}

// ---------------------------------------------------------------------------------------------
// K4 backward: red[n][c] = (sum g, sum g*xhat), g = dy * act'(y)
// ---------------------------------------------------------------------------------------------
template <typename T, typename TG, int VW, bool ST>
__global__ void __launch_bounds__(TB, 2) inorm_bwd_reduce_kernel(const TG* __restrict__ dy, const T* __restrict__ x,
                                                             const float* __restrict__ stats,
                                                             const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, int64_t voxels,
                                                             int C, float slope, int64_t rows_per_slab,
                                                             ReduceScratch rs, float* __restrict__ red,
                                                             float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                             int stages) {
  static_assert(sizeof(T) == sizeof(TG), "value and gradient storage must have the same width");
  extern __shared__ __align__(16) uint8_t dsm[];
  const int n = blockIdx.y;
  const void* const src[2] = {x + (int64_t)n * voxels * C, dy + (int64_t)n * voxels * C};
  float* out = red + (int64_t)n * C * 2;
  auto prep = [&](int cbase) { return norm_regs<VW>(stats + (int64_t)n * C * 2, gamma, beta, cbase); };
  auto body = [&](int64_t, int, const NormRegs& q, const RawVec<sizeof(T), VW> (&raw)[2], float (&acc)[2][8]) {
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
  };
  auto fin = [&](int c, const float (&t)[2]) {
    out[2 * c] = t[0];
    out[2 * c + 1] = t[1];
    // dbeta += sum g, dgamma += sum g * xhat: one atomic per (sample, channel) - no extra launch
    if (dbeta) atomicAdd(dbeta + c, t[0]);
    if (dgamma) atomicAdd(dgamma + c, t[1]);
  };
  if constexpr (ST)
    reduce_rows_staged<2, VW, 2, sizeof(T)>(src, voxels, C, rows_per_slab, stages, dsm, rs, prep, body, fin);
  else
    reduce_rows<2, VW, 4, 2, sizeof(T)>(src, voxels, C, rows_per_slab, reinterpret_cast<float*>(dsm), rs, prep, body, fin);
}

struct NormBwdRegs {
  NormRegs q;
  float c1[8], c2[8];      // (sum g) / V, (sum g * xhat) / V
};

// NTS = 2: (x, dy); NTS = 3: (x, dy, previous dx) - the accumulating launch streams the old gradient through the
// ring as a third tensor instead of a dependent global load in the loop (staged variant only)
template <typename T, typename TG, int VW, bool ST, int NTS>
__global__ void __launch_bounds__(TB, 2) inorm_bwd_apply_kernel(const TG* __restrict__ dy, const T* __restrict__ x,
                                                            const float* __restrict__ stats,
                                                            const float* __restrict__ gamma,
                                                            const float* __restrict__ beta,
                                                            const float* __restrict__ red, int64_t voxels, int C,
                                                            float slope, float inv_v, TG* __restrict__ dx,
                                                            int accumulate, int64_t rows_per_slab, int stages) {
  static_assert(ST || NTS == 2, "the register skeleton reads the previous gradient directly");
  extern __shared__ __align__(16) uint8_t dsm[];
  const int n = blockIdx.y;
  const int64_t base = (int64_t)n * voxels * C;
  const void* const src[3] = {x + base, dy + base, dx + base};
  TG* dxb = dx + base;
  auto prep = [&](int cbase) {
    NormBwdRegs w;
    w.q = norm_regs<VW>(stats + (int64_t)n * C * 2, gamma, beta, cbase);
    const float* rd = red + ((int64_t)n * C + cbase) * 2;
#pragma unroll
    for (int k = 0; k < VW; ++k) { w.c1[k] = rd[2 * k] * inv_v; w.c2[k] = rd[2 * k + 1] * inv_v; }
    return w;
  };
  auto body = [&](int64_t, int64_t off, int, const NormBwdRegs& w, const RawVec<sizeof(T), VW> (&raw)[NTS]) {
    float v[8], d[8], o[8];
    raw[0].template unpack<T>(v);
    raw[1].template unpack<TG>(d);
    if constexpr (NTS == 3) raw[2].template unpack<TG>(o);
    else if (accumulate) ldv<TG, VW>(dxb + off, o);
#pragma unroll
    for (int k = 0; k < VW; ++k) {
      const float xh = (v[k] - w.q.mean[k]) * w.q.rstd[k];
      const float gg = d[k] * (fmaf(v[k], w.q.a[k], w.q.b[k]) > 0.f ? 1.f : slope);
      const float t = w.q.a[k] * (gg - w.c1[k] - xh * w.c2[k]);
      o[k] = (NTS == 3 || accumulate) ? o[k] + t : t;
    }
    stv<TG, VW>(dxb + off, o);
  };
  if constexpr (ST && NTS == 3) {
    stream_rows_staged<VW, 3, sizeof(T)>(src, voxels, C, rows_per_slab, stages, dsm, prep, body);
  } else {
    const void* const s2[2] = {src[0], src[1]};
    if constexpr (ST) stream_rows_staged<VW, 2, sizeof(T)>(s2, voxels, C, rows_per_slab, stages, dsm, prep, body);
    else stream_rows<VW, 4, 2, sizeof(T)>(s2, voxels, C, rows_per_slab, prep, body);
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

// one block per sample; Cr <= TB. The two small matrix-vector products are spread over the whole block:
// conv6 as (TB / Cr) partial sums per hidden unit (coalesced along j) folded in a fixed order, conv7 one thread per
// channel (coalesced along c).
__global__ void __launch_bounds__(TB) se_excite_fwd_kernel(float* __restrict__ pool, const float* __restrict__ w6,
                                                          const float* __restrict__ b6, const float* __restrict__ w7,
                                                          const float* __restrict__ b7, int C, int Cr,
                                                          float* __restrict__ hidden, float* __restrict__ gate,
                                                          const float* __restrict__ stats3,
                                                          const float* __restrict__ gamma3,
                                                          const float* __restrict__ beta3) {
  extern __shared__ float sm[];  // pool[C] | act[Cr] | partial[TB]
  float* sp = sm;
  float* sa = sm + C;
  float* part = sa + Cr;
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
  const int nparts = TB / Cr;
  if ((int)threadIdx.x < nparts * Cr) {
    const int j = threadIdx.x % Cr, q = threadIdx.x / Cr;
    float h = 0.f;
    for (int c = q; c < C; c += nparts) h = fmaf(sp[c], w6[(int64_t)c * Cr + j], h);
    part[q * Cr + j] = h;
  }
  __syncthreads();
  for (int j = threadIdx.x; j < Cr; j += TB) {
    float h = b6[j];
    for (int q = 0; q < nparts; ++q) h += part[q * Cr + j];
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

// Backward of the excite MLP in two launches, no atomics (deterministic, and ~C*Cr atomics per sample were the
// whole cost of the old single kernel):
//   A  one block per sample: d7 = dgate * g(1-g), dh = (W7 d7) * lrelu'(hidden), dpool = W6 dh;
//      d7 / dh go to the context scratch [batch][C] | [batch][Cr]
//   B  one thread per weight element: dW7 += sum_n act_n d7_n^T, dW6 += sum_n pool_n dh_n^T, the bias gradients
//      and the norm3 / norm4 parameter gradients from the reductions of the gate backward (red5), summed over the
//      samples in order
__global__ void __launch_bounds__(TB) se_excite_bwd_sample_kernel(const float* __restrict__ dgate,
                                                                  const float* __restrict__ hidden,
                                                                  const float* __restrict__ gate,
                                                                  const float* __restrict__ w6,
                                                                  const float* __restrict__ w7, int C, int Cr,
                                                                  float* __restrict__ dpool, float* __restrict__ d7,
                                                                  float* __restrict__ dh) {
  extern __shared__ float sm[];  // dpre7[C] | dhid[Cr]
  float* sd7 = sm;
  float* sdh = sm + C;
  const int n = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += TB) {
    const float g = gate[(int64_t)n * C + c];
    const float d = dgate[(int64_t)n * C + c] * g * (1.f - g);
    sd7[c] = d;
    d7[(int64_t)n * C + c] = d;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int j = warp; j < Cr; j += TB / 32) {            // a warp per hidden unit: coalesced along c
    float da = 0.f;
    for (int c = lane; c < C; c += 32) da = fmaf(w7[(int64_t)j * C + c], sd7[c], da);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) da += __shfl_xor_sync(0xffffffffu, da, o);
    if (lane == 0) {
      const float v = da * (hidden[(int64_t)n * Cr + j] > 0.f ? 1.f : M1_LRELU_SLOPE);
      sdh[j] = v;
      dh[(int64_t)n * Cr + j] = v;
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += TB) {
    float dp = 0.f;
    for (int j = 0; j < Cr; ++j) dp = fmaf(w6[(int64_t)c * Cr + j], sdh[j], dp);
    dpool[(int64_t)n * C + c] = dp;
  }
}

__global__ void __launch_bounds__(TB) se_excite_bwd_param_kernel(const float* __restrict__ d7, const float* __restrict__ dh,
                                                                 const float* __restrict__ pool,
                                                                 const float* __restrict__ hidden,
                                                                 const float* __restrict__ dpool, int batch, int C,
                                                                 int Cr, float* __restrict__ dw6, float* __restrict__ db6,
                                                                 float* __restrict__ dw7, float* __restrict__ db7,
                                                                 const float* __restrict__ red5, float* __restrict__ dgamma3,
                                                                 float* __restrict__ dbeta3, float* __restrict__ dgamma4,
                                                                 float* __restrict__ dbeta4) {
  const int i = blockIdx.x * TB + threadIdx.x;
  if (i < C * Cr) {
    const int j7 = i / C, c7 = i % C;           // dW7[j][c] += act[j] * dpre7[c]
    const int c6 = i / Cr, j6 = i % Cr;         // dW6[c][j] += pool[c] * dhid[j]
    float s7 = 0.f, s6 = 0.f;
    for (int n = 0; n < batch; ++n) {
      s7 = fmaf(lrelu(hidden[(int64_t)n * Cr + j7], M1_LRELU_SLOPE), d7[(int64_t)n * C + c7], s7);
      s6 = fmaf(pool[(int64_t)n * C + c6], dh[(int64_t)n * Cr + j6], s6);
    }
    dw7[i] += s7;
    dw6[i] += s6;
  }
  if (i < C) {
    float s = 0.f;
    for (int n = 0; n < batch; ++n) s += d7[(int64_t)n * C + i];
    db7[i] += s;
    if (red5 != nullptr) {
      // norm3: dgamma += A2, dbeta += A1 + dpool ; norm4: dgamma += B2, dbeta += B1 (reductions of the gate backward)
      float g3 = 0.f, b3 = 0.f, g4 = 0.f, b4 = 0.f;
      for (int n = 0; n < batch; ++n) {
        const float* r = red5 + ((int64_t)n * C + i) * 5;
        g3 += r[1];
        b3 += r[0] + dpool[(int64_t)n * C + i];
        g4 += r[3];
        b4 += r[2];
      }
      dgamma3[i] += g3;
      dbeta3[i] += b3;
      dgamma4[i] += g4;
      dbeta4[i] += b4;
    }
  }
  if (i < Cr) {
    float s = 0.f;
    for (int n = 0; n < batch; ++n) s += dh[(int64_t)n * Cr + i];
    db6[i] += s;
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

// Dropout source of a launch, a template parameter of the gate kernels (the three code paths inlined per row made
// them 3 500 instructions long): 0 no dropout, 1 keep-bits read from m1_dropout.mask (backward), 2 injected
// uniforms, 3 Philox (the forward kernel also leaves the keep-bits in m1_dropout.mask if given)
// 4 = decided at run time from the m1_dropout fields: the FORWARD gate kernel keeps this form - measured 25 % faster
// than its Philox-only instantiation (4.4 vs 3.4 TB/s at 8 x 20x160x160 x 32), while both backward kernels gain
// 9-14 % from the specialisation.
enum { DROP_NONE = 0, DROP_MASK = 1, DROP_U = 2, DROP_PHILOX = 3, DROP_DYN = 4 };
inline int drop_mode(const DropArgs& dr, bool backward, int vw) {
  if (dr.rate <= 0.f) return DROP_NONE;
  if (backward && vw >= 4 && dr.mask != nullptr) return DROP_MASK;
  return dr.u != nullptr ? DROP_U : DROP_PHILOX;
}

// keep-mask * scale for the VW elements starting at flat element index e (e % VW == 0)
template <int VW, int DM, bool kBackward = false>
__device__ __forceinline__ void drop_factors(const DropArgs& dr, int64_t e, float (&f)[8]) {
  if constexpr (DM == DROP_DYN) {
    if (dr.rate <= 0.f) {
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = 1.f;
      return;
    }
    if constexpr (kBackward && VW >= 4) {
      if (dr.mask != nullptr) { drop_factors<VW, DROP_MASK, kBackward>(dr, e, f); return; }
    }
    if (dr.u != nullptr) drop_factors<VW, DROP_U, kBackward>(dr, e, f);
    else drop_factors<VW, DROP_PHILOX, kBackward>(dr, e, f);
    return;
  } else if constexpr (DM == DROP_NONE) {
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = 1.f;
    return;
  } else if constexpr (DM == DROP_MASK) {
    const uint32_t bits = (uint32_t)dr.mask[e >> 3] >> (e & 7);      // e % VW == 0
#pragma unroll
    for (int i = 0; i < VW; ++i) f[i] = ((bits >> i) & 1u) ? dr.scale : 0.f;
    return;
  } else {
    float u[8];
    if constexpr (DM == DROP_U) {
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
        const bool hi = ((e >> 2) & 1) != 0;                         // second half of the block: words 2, 3
        const uint32_t w0 = hi ? w[2] : w[0], w1 = hi ? w[3] : w[1];
        u[0] = (float)(w0 & 0xFFFFu) * (1.0f / 65536.0f);
        u[1] = (float)(w0 >> 16) * (1.0f / 65536.0f);
        u[2] = (float)(w1 & 0xFFFFu) * (1.0f / 65536.0f);
        u[3] = (float)(w1 >> 16) * (1.0f / 65536.0f);
      } else {
        const int q = (int)((e >> 1) & 3);
        const uint32_t ww = q == 0 ? w[0] : q == 1 ? w[1] : q == 2 ? w[2] : w[3];
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
}

constexpr bool kGatePipe = true;      // software-pipelined loads in the three gate kernels (see pipelined_rows)

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

template <typename T, int VW, bool ST, int DM>
__global__ void __launch_bounds__(TB, 2) se_gate_fwd_kernel(const T* __restrict__ raw3, const T* __restrict__ raw4,
                                                        GateArgs a, DropArgs dr, int64_t voxels, int C,
                                                        T* __restrict__ out, __nv_bfloat16* __restrict__ out2,
                                                        int64_t rows_per_slab, int stages) {
  extern __shared__ __align__(16) uint8_t dsm[];
  const int n = blockIdx.y;
  const int64_t base = (int64_t)n * voxels * C;
  const void* const src[2] = {raw3 + base, raw4 + base};
  auto prep = [&](int cbase) { return gate_regs<VW>(a, n, C, cbase); };
  auto body = [&](int64_t, int64_t off, int, const GateRegs& w, const RawVec<sizeof(T), VW> (&raw)[2]) {
    const int64_t e = base + off;
    float x3[8], x4[8], f[8], o[8];
    raw[0].template unpack<T>(x3);
    raw[1].template unpack<T>(x4);
    drop_factors<VW, DM>(dr, e, f);
#pragma unroll
    for (int k = 0; k < VW; ++k) {
      const float x_ = fmaf(x3[k], w.q3.a[k], w.q3.b[k]);
      const float res = fmaf(x4[k], w.q4.a[k], w.q4.b[k]);
      o[k] = lrelu(x_ * w.gt[k] * res, M1_LRELU_SLOPE) * f[k];
    }
    stv<T, VW>(out + e, o);
    if (out2) stv<__nv_bfloat16, VW>(out2 + e, o);
  };
  if constexpr (ST) stream_rows_staged<VW, 2, sizeof(T)>(src, voxels, C, rows_per_slab, stages, dsm, prep, body);
  else stream_rows<VW, 2, 2, sizeof(T), kGatePipe>(src, voxels, C, rows_per_slab, prep, body);
}

// red[n][c] = { sum dx_, sum dx_*xh3, sum dr, sum dr*xh4, sum dz*x_*r } ; dgate[n][c] = the last one
template <typename T, typename TG, int VW, bool ST, int DM>
__global__ void __launch_bounds__(TB, 2) se_gate_bwd_reduce_kernel(const TG* __restrict__ dout, const T* __restrict__ raw3,
                                                               const T* __restrict__ raw4, GateArgs a, DropArgs dr,
                                                               int64_t voxels, int C, int64_t rows_per_slab,
                                                               ReduceScratch rs, float* __restrict__ red5,
                                                               float* __restrict__ dgate, int stages) {
  static_assert(sizeof(T) == sizeof(TG), "value and gradient storage must have the same width");
  extern __shared__ __align__(16) uint8_t dsm[];
  const int n = blockIdx.y;
  const int64_t base = (int64_t)n * voxels * C;
  const void* const src[3] = {raw3 + base, raw4 + base, dout + base};
  float* out = red5 + (int64_t)n * C * 5;
  float* dg = dgate + (int64_t)n * C;
  auto prep = [&](int cbase) { return gate_regs<VW>(a, n, C, cbase); };
  auto body = [&](int64_t r, int cbase, const GateRegs& w, const RawVec<sizeof(T), VW> (&raw)[3], float (&acc)[5][8]) {
    const int64_t e = base + r * C + cbase;
    float x3[8], x4[8], d[8], f[8];
    raw[0].template unpack<T>(x3);
    raw[1].template unpack<T>(x4);
    raw[2].template unpack<TG>(d);
    drop_factors<VW, DM, true>(dr, e, f);
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
  };
  auto fin = [&](int c, const float (&t)[5]) {
#pragma unroll
    for (int k = 0; k < 5; ++k) out[5 * c + k] = t[k];
    dg[c] = t[4];
  };
  if constexpr (ST)
    reduce_rows_staged<5, VW, 3, sizeof(T)>(src, voxels, C, rows_per_slab, stages, dsm, rs, prep, body, fin);
  else
    reduce_rows<5, VW, 2, 3, sizeof(T), kGatePipe>(src, voxels, C, rows_per_slab, reinterpret_cast<float*>(dsm), rs, prep, body,
                                                    fin);
}

struct GateBwdRegs {
  GateRegs w;
  float c[4][8];      // the four reduction results / V
};

// accumulate (raw3 / raw4 shared by several gate kernels - one trunk, several dropouts): draw3 / draw4 +=
template <typename T, typename TG, int VW, bool ST, int DM>
__global__ void __launch_bounds__(TB, 2) se_gate_bwd_apply_kernel(const TG* __restrict__ dout, const T* __restrict__ raw3,
                                                              const T* __restrict__ raw4, GateArgs a, DropArgs dr,
                                                              const float* __restrict__ red5, int64_t voxels, int C,
                                                              float inv_v, TG* __restrict__ draw3,
                                                              TG* __restrict__ draw4, int accumulate,
                                                              int64_t rows_per_slab, int stages) {
  extern __shared__ __align__(16) uint8_t dsm[];
  const int n = blockIdx.y;
  const int64_t base = (int64_t)n * voxels * C;
  const void* const src[3] = {raw3 + base, raw4 + base, dout + base};
  auto prep = [&](int cbase) {
    GateBwdRegs g;
    g.w = gate_regs<VW>(a, n, C, cbase);
    const float* rd = red5 + ((int64_t)n * C + cbase) * 5;
#pragma unroll
    for (int k = 0; k < VW; ++k)
#pragma unroll
      for (int j = 0; j < 4; ++j) g.c[j][k] = rd[5 * k + j] * inv_v;
    return g;
  };
  auto body = [&](int64_t, int64_t off, int, const GateBwdRegs& g, const RawVec<sizeof(T), VW> (&raw)[3]) {
    const GateRegs& w = g.w;
    const int64_t e = base + off;
    float x3[8], x4[8], d[8], f[8], o3[8], o4[8];
    raw[0].template unpack<T>(x3);
    raw[1].template unpack<T>(x4);
    raw[2].template unpack<TG>(d);
    drop_factors<VW, DM, true>(dr, e, f);
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
    if (accumulate) {
      float p3[8], p4[8];
      ldv<TG, VW>(draw3 + e, p3);
      ldv<TG, VW>(draw4 + e, p4);
#pragma unroll
      for (int k = 0; k < VW; ++k) { o3[k] += p3[k]; o4[k] += p4[k]; }
    }
    stv<TG, VW>(draw3 + e, o3);
    stv<TG, VW>(draw4 + e, o4);
  };
  if constexpr (ST) stream_rows_staged<VW, 3, sizeof(T)>(src, voxels, C, rows_per_slab, stages, dsm, prep, body);
  else stream_rows<VW, 2, 3, sizeof(T), kGatePipe>(src, voxels, C, rows_per_slab, prep, body);
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

// ---- staged launches: eligibility, ring depth, slab size ------------------------------------------------------
struct StagedPlan {
  bool on;
  int stages;
  size_t smem;
  int64_t rows;       // rows per slab: a multiple of the rows of one ring stage
};
// `which`: 1 inorm_stats, 2 inorm_act_fwd, 4 inorm_act_bwd. Measured on B200 (tools/bench_elementwise.py, 8 x 20x160x160
// x 32 fp16): statistics 3.4 staged / 4.0 TB/s register skeleton, forward apply 5.5 / 5.4, backward pair 5.6 / 4.6 -
// the default (M1_STAGED unset) stages the backward pair only.
inline StagedPlan staged_plan(const m1_ctx* ctx, int which, int batch, int64_t voxels, int C, int es, int vw, int nt,
                              int per_sm = 4) {
  static const int enabled = getenv("M1_STAGED") ? atoi(getenv("M1_STAGED")) : 4;
  StagedPlan sp{false, 0, 0, 0};
  if (!(enabled & which) || vw < 4 || C % vw) return sp;
  const int cg = C / vw;
  if (((int64_t)C * es) % 16 || cg > TB || TB % cg) return sp;
  const int m = kStageBytes / (TB * vw * es);
  if (m < 1) return sp;
  const int rows_chunk = m * (TB / cg);
  sp.stages = std::min(kMaxStages, (96 * 1024) / (nt * kStageBytes));       // NT 3: 4 stages, 2: 6, 1: 8
  if (sp.stages < 3) return sp;
  const int64_t slabs = std::max<int64_t>(1, (int64_t)ctx->num_sms * per_sm / std::max(1, batch));
  int64_t rows = cdiv64(voxels, slabs);
  rows = cdiv64(std::max<int64_t>(rows, 1), rows_chunk) * rows_chunk;
  sp.rows = rows;
  sp.smem = 256 + (size_t)sp.stages * nt * kStageBytes;      // 128 B of barriers + alignment slack
  sp.on = true;
  return sp;
}
// > 48 KB of dynamic shared memory needs the opt-in once per kernel
template <typename K>
inline int allow_smem(K kernel, size_t bytes) {
  static std::unordered_set<const void*> done;
  const void* key = reinterpret_cast<const void*>(kernel);
  if (bytes > 48 * 1024 && !done.count(key)) {
    M1_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    done.insert(key);
  }
  return 0;
}
inline int es_of(int dtype) { return dtype == M1_F32 ? 4 : 2; }

template <typename T, int VW>
int launch_inorm_stats(m1_ctx* ctx, const T* x, int batch, int64_t voxels, int C, float eps, float* stats,
                       cudaStream_t st) {
  ReduceScratch rs;
  if constexpr (VW > 1) {
    const StagedPlan sp = staged_plan(ctx, 1, batch, voxels, C, sizeof(T), VW, 1);
    if (sp.on) {
      dim3 grid((unsigned)cdiv64(voxels, sp.rows), (unsigned)batch);
      if (reduce_scratch(ctx, grid.x, batch, 2, C, &rs)) return 1;
      if (allow_smem(inorm_stats_kernel<T, VW, true>, sp.smem)) return 1;
      inorm_stats_kernel<T, VW, true><<<grid, TB, sp.smem, st>>>(x, voxels, C, sp.rows, 1.f / (float)voxels, eps, rs, stats,
                                                                 sp.stages);
      return 0;
    }
  }
  const int64_t rows = slab_rows(ctx, batch, voxels);
  dim3 grid((unsigned)cdiv64(voxels, rows), (unsigned)batch);
  if (reduce_scratch(ctx, grid.x, batch, 2, C, &rs)) return 1;
  inorm_stats_kernel<T, VW, false><<<grid, TB, reduce_smem(2, C, VW), st>>>(x, voxels, C, rows, 1.f / (float)voxels, eps, rs,
                                                                            stats, 0);
  return 0;
}

template <typename T, int VW>
int launch_inorm_act_fwd(m1_ctx* ctx, const T* x, const float* stats, const float* gamma, const float* beta, int batch,
                         int64_t voxels, int C, float slope, T* y, __nv_bfloat16* y2, cudaStream_t st) {
  if constexpr (VW > 1) {
    const StagedPlan sp = staged_plan(ctx, 2, batch, voxels, C, sizeof(T), VW, 1);
    if (sp.on) {
      dim3 grid((unsigned)cdiv64(voxels, sp.rows), (unsigned)batch);
      if (allow_smem(inorm_act_fwd_kernel<T, VW, true>, sp.smem)) return 1;
      inorm_act_fwd_kernel<T, VW, true><<<grid, TB, sp.smem, st>>>(x, stats, gamma, beta, voxels, C, slope, y, y2, sp.rows,
                                                                   sp.stages);
      return 0;
    }
  }
  const int64_t rows = slab_rows(ctx, batch, voxels, 8);
  dim3 grid((unsigned)cdiv64(voxels, rows), (unsigned)batch);
  inorm_act_fwd_kernel<T, VW, false><<<grid, TB, 0, st>>>(x, stats, gamma, beta, voxels, C, slope, y, y2, rows, 0);
  return 0;
}

template <typename T, typename TG, int VW>
int launch_inorm_act_bwd(m1_ctx* ctx, const TG* dy, const T* x, const float* stats, const float* gamma, const float* beta,
                         int batch, int64_t voxels, int C, float slope, TG* dx, int accumulate, float* dgamma,
                         float* dbeta, float* red, cudaStream_t st) {
  ReduceScratch rs;
  const float inv_v = 1.f / (float)voxels;
  if constexpr (VW > 1) {
    const StagedPlan sp = staged_plan(ctx, 4, batch, voxels, C, sizeof(T), VW, 2);
    const StagedPlan sa = staged_plan(ctx, 4, batch, voxels, C, sizeof(T), VW, accumulate ? 3 : 2);
    if (sp.on && sa.on) {
      dim3 grid((unsigned)cdiv64(voxels, sp.rows), (unsigned)batch);
      if (reduce_scratch(ctx, grid.x, batch, 2, C, &rs)) return 1;
      if (allow_smem(inorm_bwd_reduce_kernel<T, TG, VW, true>, sp.smem)) return 1;
      inorm_bwd_reduce_kernel<T, TG, VW, true><<<grid, TB, sp.smem, st>>>(dy, x, stats, gamma, beta, voxels, C, slope, sp.rows,
                                                                          rs, red, dgamma, dbeta, sp.stages);
      M1_LAUNCH_CHECK(ctx);
      dim3 grid2((unsigned)cdiv64(voxels, sa.rows), (unsigned)batch);
      if (accumulate) {
        if (allow_smem(inorm_bwd_apply_kernel<T, TG, VW, true, 3>, sa.smem)) return 1;
        inorm_bwd_apply_kernel<T, TG, VW, true, 3><<<grid2, TB, sa.smem, st>>>(dy, x, stats, gamma, beta, red, voxels, C, slope,
                                                                               inv_v, dx, 1, sa.rows, sa.stages);
      } else {
        if (allow_smem(inorm_bwd_apply_kernel<T, TG, VW, true, 2>, sa.smem)) return 1;
        inorm_bwd_apply_kernel<T, TG, VW, true, 2><<<grid2, TB, sa.smem, st>>>(dy, x, stats, gamma, beta, red, voxels, C, slope,
                                                                               inv_v, dx, 0, sa.rows, sa.stages);
      }
      return 0;
    }
  }
  const int64_t rows = slab_rows(ctx, batch, voxels);
  dim3 grid((unsigned)cdiv64(voxels, rows), (unsigned)batch);
  if (reduce_scratch(ctx, grid.x, batch, 2, C, &rs)) return 1;
  inorm_bwd_reduce_kernel<T, TG, VW, false><<<grid, TB, reduce_smem(2, C, VW), st>>>(dy, x, stats, gamma, beta, voxels, C, slope,
                                                                                     rows, rs, red, dgamma, dbeta, 0);
  M1_LAUNCH_CHECK(ctx);
  const int64_t rows8 = slab_rows(ctx, batch, voxels, 8);
  dim3 grid8((unsigned)cdiv64(voxels, rows8), (unsigned)batch);
  inorm_bwd_apply_kernel<T, TG, VW, false, 2><<<grid8, TB, 0, st>>>(dy, x, stats, gamma, beta, red, voxels, C, slope, inv_v, dx,
                                                                    accumulate, rows8, 0);
  return 0;
}

// The gate kernels are instruction-bound (20+ FP32 operations per channel and tensor pass), not latency-bound: the
// staged skeleton buys them nothing (measured: 2.2 TB/s either way for the backward reduction), so they keep the
// register skeleton and are specialised on the dropout source instead.
#define M1_DISPATCH_DM(mode, ...)                                      \
  do {                                                                 \
    switch (mode) {                                                    \
      case DROP_NONE: { constexpr int DM = DROP_NONE; __VA_ARGS__; } break;     \
      case DROP_MASK: { constexpr int DM = DROP_MASK; __VA_ARGS__; } break;     \
      case DROP_U: { constexpr int DM = DROP_U; __VA_ARGS__; } break;           \
      default: { constexpr int DM = DROP_PHILOX; __VA_ARGS__; } break;          \
    }                                                                  \
  } while (0)

template <typename T, int VW>
int launch_se_gate_fwd(m1_ctx* ctx, const T* raw3, const T* raw4, const GateArgs& a, const DropArgs& dr, int batch,
                       int64_t voxels, int C, T* out, __nv_bfloat16* out2, cudaStream_t st) {
  const int64_t rows = slab_rows(ctx, batch, voxels, 8);
  dim3 grid((unsigned)cdiv64(voxels, rows), (unsigned)batch);
  se_gate_fwd_kernel<T, VW, false, DROP_DYN><<<grid, TB, 0, st>>>(raw3, raw4, a, dr, voxels, C, out, out2, rows, 0);
  return 0;
}

template <typename T, typename TG, int VW>
int launch_se_gate_bwd_reduce(m1_ctx* ctx, const TG* dout, const T* raw3, const T* raw4, const GateArgs& a,
                              const DropArgs& dr, int batch, int64_t voxels, int C, float* red, float* dgate,
                              cudaStream_t st) {
  ReduceScratch rs;
  const int64_t rows = slab_rows(ctx, batch, voxels);
  dim3 grid((unsigned)cdiv64(voxels, rows), (unsigned)batch);
  if (reduce_scratch(ctx, grid.x, batch, 5, C, &rs)) return 1;
  M1_DISPATCH_DM(drop_mode(dr, true, VW), (se_gate_bwd_reduce_kernel<T, TG, VW, false, DM><<<grid, TB, reduce_smem(5, C, VW), st>>>(
                                              dout, raw3, raw4, a, dr, voxels, C, rows, rs, red, dgate, 0)));
  return 0;
}

template <typename T, typename TG, int VW>
int launch_se_gate_bwd_apply(m1_ctx* ctx, const TG* dout, const T* raw3, const T* raw4, const GateArgs& a,
                             const DropArgs& dr, const float* red, int batch, int64_t voxels, int C, TG* draw3, TG* draw4,
                             int accumulate, cudaStream_t st) {
  const float inv_v = 1.f / (float)voxels;
  const int64_t rows = slab_rows(ctx, batch, voxels, 8);
  dim3 grid((unsigned)cdiv64(voxels, rows), (unsigned)batch);
  M1_DISPATCH_DM(drop_mode(dr, true, VW), (se_gate_bwd_apply_kernel<T, TG, VW, false, DM><<<grid, TB, 0, st>>>(
                                              dout, raw3, raw4, a, dr, red, voxels, C, inv_v, draw3, draw4, accumulate, rows, 0)));
  return 0;
}

}  // namespace

extern "C" int m1_inorm_stats(m1_ctx* ctx, const void* x, int dtype, int batch, int64_t voxels, int C,
                              float eps, float* stats, void* stream) {
  int rc = 0;
  DISPATCH_T_VW(dtype, C, (rc = launch_inorm_stats<T, VW>(ctx, reinterpret_cast<const T*>(x), batch, voxels, C, eps, stats,
                                                         (cudaStream_t)stream)));
  if (rc) return rc;
  M1_LAUNCH_CHECK(ctx);
  return 0;
}

extern "C" int m1_inorm_act_fwd(m1_ctx* ctx, const void* x, const float* stats, const float* gamma,
                                const float* beta, int dtype, int batch, int64_t voxels, int C, float slope,
                                void* y, void* y_bf16, void* stream) {
  int rc = 0;
  DISPATCH_T_VW(dtype, C, (rc = launch_inorm_act_fwd<T, VW>(ctx, reinterpret_cast<const T*>(x), stats, gamma, beta, batch, voxels,
                                                           C, slope, reinterpret_cast<T*>(y),
                                                           reinterpret_cast<__nv_bfloat16*>(y_bf16), (cudaStream_t)stream)));
  if (rc) return rc;
  M1_LAUNCH_CHECK(ctx);
  return 0;
}

extern "C" int m1_inorm_act_bwd(m1_ctx* ctx, const void* dy, const void* x, const float* stats,
                                const float* gamma, const float* beta, int dtype, int batch, int64_t voxels,
                                int C, float slope, void* dx, int accumulate, float* dgamma, float* dbeta,
                                void* stream) {
  M1_CHECK((size_t)batch * C * 2 * sizeof(float) <= ctx->scratch_bytes, "inorm_act_bwd: scratch too small");
  int rc = 0;
  DISPATCH_T_VW4(dtype, C, (rc = launch_inorm_act_bwd<T, TG, VW>(ctx, reinterpret_cast<const TG*>(dy),
                                                                reinterpret_cast<const T*>(x), stats, gamma, beta, batch,
                                                                voxels, C, slope, reinterpret_cast<TG*>(dx), accumulate,
                                                                dgamma, dbeta, ctx->scratch, (cudaStream_t)stream)));
  if (rc) return rc;
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
  M1_CHECK(Cr >= 1 && Cr <= TB, "m1_se_excite_fwd: Cr = %d outside [1, %d]", Cr, TB);
  se_excite_fwd_kernel<<<batch, TB, (C + Cr + TB) * sizeof(float), (cudaStream_t)stream>>>(
      pool, w6, b6, w7, b7, C, Cr, hidden, gate, stats3, gamma3, beta3);
  M1_LAUNCH_CHECK(ctx);
  return 0;
}

extern "C" int m1_se_excite_bwd(m1_ctx* ctx, const float* dgate, const float* pool, const float* hidden,
                                const float* gate, const float* w6, const float* w7, int batch, int C, int Cr,
                                float* dpool, float* dw6, float* db6, float* dw7, float* db7, const float* red5,
                                float* dgamma3, float* dbeta3, float* dgamma4, float* dbeta4, void* stream) {
  M1_CHECK(red5 == nullptr || (dgamma3 && dbeta3 && dgamma4 && dbeta4),
           "m1_se_excite_bwd: red5 given without the four norm parameter gradients");
  M1_CHECK((size_t)batch * (C + Cr) * sizeof(float) <= ctx->scratch_bytes, "m1_se_excite_bwd: scratch too small");
  cudaStream_t st = (cudaStream_t)stream;
  float* d7 = ctx->scratch;
  float* dh = d7 + (size_t)batch * C;
  se_excite_bwd_sample_kernel<<<batch, TB, (C + Cr) * sizeof(float), st>>>(dgate, hidden, gate, w6, w7, C, Cr, dpool,
                                                                           d7, dh);
  M1_LAUNCH_CHECK(ctx);
  se_excite_bwd_param_kernel<<<(C * Cr + TB - 1) / TB, TB, 0, st>>>(d7, dh, pool, hidden, dpool, batch, C, Cr, dw6, db6,
                                                                   dw7, db7, red5, dgamma3, dbeta3, dgamma4, dbeta4);
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
  int rc = 0;
  DISPATCH_T_VW(dtype, C, (rc = launch_se_gate_fwd<T, VW>(ctx, reinterpret_cast<const T*>(raw3), reinterpret_cast<const T*>(raw4),
                                                         a, dr, batch, voxels, C, reinterpret_cast<T*>(out),
                                                         reinterpret_cast<__nv_bfloat16*>(out_bf16), (cudaStream_t)stream)));
  if (rc) return rc;
  M1_LAUNCH_CHECK(ctx);
  return 0;
}

extern "C" int m1_se_gate_bwd_reduce(m1_ctx* ctx, const void* dout, const void* raw3, const void* raw4,
                                     const float* stats3, const float* stats4, const float* gamma3,
                                     const float* beta3, const float* gamma4, const float* beta4,
                                     const float* gate, const m1_dropout* drop, int dtype, int batch,
                                     int64_t voxels, int C, float* red, float* dgate, void* stream) {
  GateArgs a{stats3, stats4, gamma3, beta3, gamma4, beta4, gate};
  DropArgs dr = make_drop(drop);
  int rc = 0;
  DISPATCH_T_VW4(dtype, C, (rc = launch_se_gate_bwd_reduce<T, TG, VW>(ctx, reinterpret_cast<const TG*>(dout),
                                                                     reinterpret_cast<const T*>(raw3),
                                                                     reinterpret_cast<const T*>(raw4), a, dr, batch, voxels, C,
                                                                     red, dgate, (cudaStream_t)stream)));
  if (rc) return rc;
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
  int rc = 0;
  DISPATCH_T_VW4(dtype, C, (rc = launch_se_gate_bwd_apply<T, TG, VW>(ctx, reinterpret_cast<const TG*>(dout),
                                                                    reinterpret_cast<const T*>(raw3),
                                                                    reinterpret_cast<const T*>(raw4), a, dr, red, batch, voxels,
                                                                    C, reinterpret_cast<TG*>(draw3), reinterpret_cast<TG*>(draw4),
                                                                    accumulate, st)));
  if (rc) return rc;
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
