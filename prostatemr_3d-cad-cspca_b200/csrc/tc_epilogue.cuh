// tc_epilogue.cuh — shared epilogue of the tcgen05 convolution kernels: one accumulator row (an output
// voxel) per thread, TMEM -> registers -> (+bias, +previous contents) -> bf16 -> global.
//
// The epilogue is instruction-latency bound (few resident warps): the produced-tensor lookup is a table in
// the kernel parameters (one entry per 8-column group), TMEM is read 16 columns at a time, the previous
// contents of accumulating launches are fetched before the TMEM wait, and the kernels run EIGHT epilogue
// warps - warps w and w+4 share a TMEM lane quarter and split the columns.
#pragma once
#include "tc_common.cuh"

namespace tc {

constexpr int kMaxGroups = 128;   // 8-column groups over all produced channels of a launch (<= 1024 channels)

struct EpiOut {
  void* out[M1_MAX_OUT];
  const float* bias[M1_MAX_OUT];
  int out_c[M1_MAX_OUT];
  int accumulate;                 // bitmask over produced tensors
  int n_total;                    // real produced channels (multiple of 8)
  int out_f16;                    // produced tensors are fp16 (else bf16)
  uint8_t grp_out[kMaxGroups];    // 8-column group -> produced tensor
  uint16_t grp_c[kMaxGroups];     // 8-column group -> first channel inside that tensor
};

// host: fill the group tables; returns false if the launch has too many produced channels
inline bool epi_fill(EpiOut* e, const m1_conv_desc* d, const float* const* bias, void* const* outs) {
  int total = 0;
  for (int j = 0; j < d->nout; ++j) {
    if (d->out_c[j] % 8) return false;
    for (int c = 0; c < d->out_c[j]; c += 8) {
      const int g = (total + c) / 8;
      if (g >= kMaxGroups) return false;
      e->grp_out[g] = (uint8_t)j;
      e->grp_c[g] = (uint16_t)c;
    }
    total += d->out_c[j];
    e->out[j] = outs[j];
    e->out_c[j] = d->out_c[j];
    e->bias[j] = bias ? bias[j] : nullptr;
  }
  e->n_total = total;
  e->accumulate = d->accumulate;
  e->out_f16 = d->out_dtype == M1_F16;
  return true;
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// Columns [col_begin, col_end) (multiples of 16) of the accumulator row at `taddr` (TMEM address of column 0
// of this tile, lane field already set) -> produced channels n0 + column of voxel `vox`.
// All 32 lanes of the warp must call (tcgen05.ld is warp-collective); `valid` masks the stores.
__device__ __forceinline__ void epilogue_row(const EpiOut& e, uint32_t taddr, int n0, int col_begin, int col_end,
                                             bool valid, int64_t vox, bool zero_acc) {
  for (int j = col_begin; j < col_end; j += 16) {
    uint4 old[2];
    uint16_t* dst[2];
    const float* bias[2];
    bool live[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int gcol = n0 + j + 8 * h;
      live[h] = valid && gcol < e.n_total;
      const int grp = min(gcol >> 3, kMaxGroups - 1);
      const int o = e.grp_out[grp], c = e.grp_c[grp];
      dst[h] = reinterpret_cast<uint16_t*>(e.out[o]) + vox * e.out_c[o] + c;
      bias[h] = e.bias[o] ? e.bias[o] + c : nullptr;
      old[h] = make_uint4(0u, 0u, 0u, 0u);
      if (live[h] && ((e.accumulate >> o) & 1)) old[h] = *reinterpret_cast<const uint4*>(dst[h]);
    }
    uint32_t v[16];
    tmem_ld16(taddr + (uint32_t)j, v);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (!live[h]) continue;
      float f[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = zero_acc ? 0.f : __uint_as_float(v[8 * h + i]);
      if (bias[h] != nullptr) {
        const float4 b0 = *reinterpret_cast<const float4*>(bias[h]);
        const float4 b1 = *reinterpret_cast<const float4*>(bias[h] + 4);
        f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w;
        f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
      }
      const uint32_t ow[4] = {old[h].x, old[h].y, old[h].z, old[h].w};               // zeros unless accumulating
      uint32_t pk[4];
      if (e.out_f16) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 o2 = cvt2<__half>(ow[i]);
          pk[i] = pack2<__half>(f[2 * i] + o2.x, f[2 * i + 1] + o2.y);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 o2 = cvt2<__nv_bfloat16>(ow[i]);
          pk[i] = pack2<__nv_bfloat16>(f[2 * i] + o2.x, f[2 * i + 1] + o2.y);
        }
      }
      *reinterpret_cast<uint4*>(dst[h]) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
  }
}

// column range of epilogue warp `warp` (0..7) for a tile of n_tile columns: warps 0-3 take the first half
__device__ __forceinline__ void epi_cols(int warp, int n_tile, int* begin, int* end) {
  const int split = ((n_tile / 16 + 1) / 2) * 16;
  *begin = warp < 4 ? 0 : split;
  *end = warp < 4 ? split : n_tile;
}

}  // namespace tc
