// Shared device helpers of the tcgen05 attention kernels (attention_tc.cu: T = 128 / 256; attention_tc_long.cu:
// T = 512 / 1024): no-swizzle core-matrix tiles, TMEM load shapes, cp.async fill.
#pragma once
#include "common.cuh"

namespace mdt {

constexpr int kQB = 128;  // query rows per MMA (TMEM lanes) == threads per CTA

MDT_DEVINL uint64_t make_smem_desc_nosw(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3fffu);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3fffu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3fffu) << 32;
  d |= 1ull << 46;  // version = 1 (Blackwell); layout type 0 = SWIZZLE_NONE
  return d;
}
MDT_DEVINL void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
MDT_DEVINL void tmem_ld_32x32b_x8(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
// N consecutive fp32 columns (N % 8 == 0, N <= 96) of this thread's TMEM lane -> registers, ONE wait at the end
template <int N>
MDT_DEVINL void tmem_ld_cols(uint32_t taddr, uint32_t* r) {
  int c = 0;
#pragma unroll
  for (; c + 32 <= N; c += 32) tmem_ld_32x32b_x32(taddr + c, r + c);
  if (N - c >= 16) {
    tmem_ld_32x32b_x16(taddr + c, r + c);
    c += 16;
  }
  if (N - c >= 8) tmem_ld_32x32b_x8(taddr + c, r + c);
  tcgen05_wait_ld();
}
// store N fp32 values as bf16 to a global row, 16 bytes at a time, columns >= dh dropped
template <int N>
MDT_DEVINL void store_row_bf16(__nv_bfloat16* grow, int col0, const uint32_t* r, int dh, float mul) {
#pragma unroll
  for (int g = 0; g < N / 8; ++g)
    if (col0 + 8 * g < dh)
      *reinterpret_cast<uint4*>(grow + col0 + 8 * g) = make_uint4(
          pack_bf16(__uint_as_float(r[8 * g + 0]) * mul, __uint_as_float(r[8 * g + 1]) * mul),
          pack_bf16(__uint_as_float(r[8 * g + 2]) * mul, __uint_as_float(r[8 * g + 3]) * mul),
          pack_bf16(__uint_as_float(r[8 * g + 4]) * mul, __uint_as_float(r[8 * g + 5]) * mul),
          pack_bf16(__uint_as_float(r[8 * g + 6]) * mul, __uint_as_float(r[8 * g + 7]) * mul));
}
MDT_DEVINL void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}
MDT_DEVINL void sts128u(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
MDT_DEVINL uint4 ldg128u_nc(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(gaddr(p)));
  return v;
}

template <int DP>
struct TokTile {
  static constexpr int CH = DP / 8;          // 16-byte chunks per row
  static constexpr int ROWBLK = CH * 128;    // bytes per 8-row block
  static constexpr int CG4 = (CH + 3) / 4;   // chunk groups of 4
  MDT_DEVINL static uint32_t off(int row, int c8) { return (row >> 3) * ROWBLK + c8 * 128 + (row & 7) * 16; }
  // global [rows, dh] (row stride gstride elements) -> smem tile; warp item = 8 rows x 4 chunks (64 B per row from
  // global = 2 full sectors; 128 B contiguous per quarter-warp into smem = conflict free)
  // Asynchronous (cp.async, 16 B, zero-fill for the padded chunk): every chunk of every tile is in flight at once; the
  // caller does cp_async_wait_all() once before the first MMA.  (A register-staged fill was latency-bound: 11k of
  // the 22k cycles a forward CTA lived.)
  MDT_DEVINL static void load(uint32_t s_base, const __nv_bfloat16* g, long long gstride, int rows, int dh) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = lane & 7, cl = lane >> 3;
    const int items = (rows >> 3) * CG4;
    for (int it = warp; it < items; it += blockDim.x >> 5) {
      const int rb = it / CG4, cg = it - rb * CG4;
      const int c8 = cg * 4 + cl, row = rb * 8 + r;
      if (c8 < CH) {
        const bool real = c8 * 8 < dh;
        const __nv_bfloat16* src = g + row * gstride + (real ? c8 * 8 : 0);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s_base + off(row, c8)), "l"(gaddr(src)),
                     "r"(real ? 16 : 0)
                     : "memory");
      }
    }
  }
};

// N fp32 values of one row -> bf16 chunks c8_0.. of a token tile in smem (chunks at or beyond dh/8 are skipped)
template <int DP, int N>
MDT_DEVINL void stage_row_bf16(uint32_t tile, int row, int c8_0, const uint32_t* r, int dh) {
#pragma unroll
  for (int g = 0; g < N / 8; ++g)
    if ((c8_0 + g) * 8 < dh)
      sts128u(tile + TokTile<DP>::off(row, c8_0 + g),
              make_uint4(pack_bf16(__uint_as_float(r[8 * g + 0]), __uint_as_float(r[8 * g + 1])),
                         pack_bf16(__uint_as_float(r[8 * g + 2]), __uint_as_float(r[8 * g + 3])),
                         pack_bf16(__uint_as_float(r[8 * g + 4]), __uint_as_float(r[8 * g + 5])),
                         pack_bf16(__uint_as_float(r[8 * g + 6]), __uint_as_float(r[8 * g + 7]))));
}

// C[128 x N] (+)= A[128 x K] * B[N x K]^T, K-major views of token tiles; K = DP
template <int DP>
MDT_DEVINL void mma_kk(uint32_t tmem_d, uint32_t sa, uint32_t sb, int n, bool acc0) {
  constexpr uint32_t RB = TokTile<DP>::ROWBLK;
  const uint32_t idesc = make_idesc_bf16(kQB, n, 0, 0);
#pragma unroll
  for (int k = 0; k < DP / 16; ++k)
    umma_bf16(tmem_d, make_smem_desc_nosw(sa + k * 256, 128, RB), make_smem_desc_nosw(sb + k * 256, 128, RB), idesc,
              (acc0 || k > 0) ? 1u : 0u);
}

}  // namespace mdt
