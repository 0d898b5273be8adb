// Split token tiles for the TMA-fed attention kernels (attention_sw.cu, attention_sw_long.cu): operand views,
// shared-memory descriptors, the two MMA groups every kernel is built from, and the staging of gradient rows.
// Layout of a tile [rows x DP] (see the header of attention_sw.cu):
//   block A = columns 0 .. min(DP,64)-1, one TMA box, SWIZZLE_128B (128-byte rows) or SWIZZLE_64B (DP = 32);
//   block B (DP = 80) = columns 64..79 as two no-swizzle chunk planes behind block A (plane 1 = zeros).
#pragma once
#include "attention_tc.cuh"

namespace mdt {

// operand view of (a 128-row slice of) a split tile
struct SwOp {
  uint32_t a;      // block A address of the slice's first row
  uint32_t b;      // block B plane 0 address of the slice's first row
  uint32_t plane;  // bytes between the two planes of block B (= 16 * rows of the whole tile)
};
// Block A holds min(DP, 64) columns: 128-byte rows in SWIZZLE_128B atoms for head_dim 64 / 72, 64-byte rows in
// SWIZZLE_64B atoms for head_dim 32 (decoder; default since round 2, MDT_ATTN_SW64=0 disables).
constexpr int sw_row_bytes(int dp) { return dp >= 64 ? 128 : dp * 2; }
MDT_DEVINL SwOp sw_op(uint32_t tile, int tile_rows, int row0, uint32_t row_bytes = 128u) {
  return SwOp{tile + row0 * row_bytes, tile + tile_rows * row_bytes + row0 * 16u, tile_rows * 16u};
}
constexpr int sw_tile_bytes(int dp, int rows) { return rows * dp * 2; }
// descriptor of a block-A operand: SWIZZLE_128B (layout type 2) or SWIZZLE_64B (layout type 4); SBO = 8 rows
template <int DP>
MDT_DEVINL uint64_t sw_desc(uint32_t addr, uint32_t lbo) {
  if constexpr (DP >= 64) {
    return make_smem_desc_sw128(addr, lbo, 1024);
  } else {
    uint64_t d = make_smem_desc_nosw(addr, lbo, 8 * sw_row_bytes(DP));  // version bit set, layout bits clear
    return d | (4ull << 61);
  }
}

// D[128 x n] = A[128 x DP] * B[n x DP]^T   (both K-major: contraction over head_dim)
template <int DP>
MDT_DEVINL void sw_mma_kk(uint32_t tmem_d, SwOp a, SwOp b, int n) {
  const uint32_t idesc = make_idesc_bf16(kQB, n, 0, 0);
  const uint64_t da = sw_desc<DP>(a.a, 16), db = sw_desc<DP>(b.a, 16);
  constexpr int kSteps = (DP >= 64 ? 64 : DP) / 16;
#pragma unroll
  for (int k = 0; k < kSteps; ++k) umma_bf16(tmem_d, da + 2 * k, db + 2 * k, idesc, k > 0 ? 1u : 0u);  // +32 B per k-step
  if constexpr (DP > 64)
    umma_bf16(tmem_d, make_smem_desc_nosw(a.b, a.plane, 128), make_smem_desc_nosw(b.b, b.plane, 128), idesc, 1u);
}

// D[128 x DP] (+)= A[128 x 128 tokens] * B[128 tokens x DP]   (B = split tile slice, contraction over its tokens)
// a_desc0 / a_step: descriptor of A's first k-step and its increment (in 16-byte units) per 16 tokens
template <int DP>
MDT_DEVINL void sw_mma_tok(uint32_t tmem_d, uint64_t a_desc0, uint32_t a_step, int a_mn, SwOp b, bool acc0) {
  constexpr int kNA = DP >= 64 ? 64 : DP;  // output columns that come from block A
  const uint32_t i64 = make_idesc_bf16(kQB, kNA, a_mn, 1), i16 = make_idesc_bf16(kQB, 16, a_mn, 1);
  const uint64_t db = sw_desc<DP>(b.a, 8192);
  const uint64_t db2 = make_smem_desc_nosw(b.b, 128, b.plane);
#pragma unroll
  for (int k = 0; k < kQB / 16; ++k) {
    const uint64_t da = a_desc0 + static_cast<uint64_t>(k) * a_step;
    const uint32_t acc = (acc0 || k > 0) ? 1u : 0u;
    umma_bf16(tmem_d, da, db + static_cast<uint64_t>(k) * ((16 * sw_row_bytes(DP)) >> 4), i64, acc);
    if constexpr (DP > 64) umma_bf16(tmem_d + 64, da, db2 + static_cast<uint64_t>(k) * (256 >> 4), i16, acc);
  }
}

// N fp32 values of one row -> bf16 chunks c8_0.. of a 128-row split tile
template <int DP, int N>
MDT_DEVINL void stage_row_split(uint32_t tile, int row, int c8_0, const uint32_t* r, int dh) {
#pragma unroll
  for (int g = 0; g < N / 8; ++g) {
    const int c8 = c8_0 + g;
    if (c8 * 8 >= dh) continue;
    const uint32_t dst = c8 < 8 ? tile + row * 128 + ((c8 ^ (row & 7)) << 4) : tile + kQB * 128 + row * 16;
    sts128u(dst, make_uint4(pack_bf16(__uint_as_float(r[8 * g + 0]), __uint_as_float(r[8 * g + 1])),
                            pack_bf16(__uint_as_float(r[8 * g + 2]), __uint_as_float(r[8 * g + 3])),
                            pack_bf16(__uint_as_float(r[8 * g + 4]), __uint_as_float(r[8 * g + 5])),
                            pack_bf16(__uint_as_float(r[8 * g + 6]), __uint_as_float(r[8 * g + 7]))));
  }
}

// N fp32 values of row `row` -> bf16 chunks c8_0.. of a SWIZZLE_64B tile (64-byte rows, 4 chunks per row)
template <int N>
MDT_DEVINL void stage_row_sw64(uint32_t tile, int row, int c8_0, const uint32_t* r) {
#pragma unroll
  for (int g = 0; g < N / 8; ++g) {
    const int c8 = c8_0 + g;
    sts128u(tile + row * 64 + ((c8 ^ ((row >> 1) & 3)) << 4),
            make_uint4(pack_bf16(__uint_as_float(r[8 * g + 0]), __uint_as_float(r[8 * g + 1])),
                       pack_bf16(__uint_as_float(r[8 * g + 2]), __uint_as_float(r[8 * g + 3])),
                       pack_bf16(__uint_as_float(r[8 * g + 4]), __uint_as_float(r[8 * g + 5])),
                       pack_bf16(__uint_as_float(r[8 * g + 6]), __uint_as_float(r[8 * g + 7]))));
  }
}

}  // namespace mdt
