// Attention forward for head_dim 64 / 72 with TMA-friendly operand tiles.
//
// The no-swizzle core-matrix tiles of attention_tc.cu can only be filled 16 bytes at a time (cp.async, or TMA boxes
// whose inner extent is 16 bytes): r01 phase timing showed both paths limited to ~16 B/clk per SM, i.e. the tile fill
// (and, in the backward, the gradient write-out) cost as much as all the math of an item.  Here a token tile
// [rows x DP] is split by columns:
//   block A: columns 0..63, 128 bytes per row, SWIZZLE_128B atoms (8 rows x 128 B)  - ONE TMA box of 64 x rows
//            elements (full 128-byte bursts); the same bytes serve as a K-major operand (contraction over head_dim,
//            32 bytes per k-step inside the atom) and as an MN-major operand (contraction over tokens, N = 64,
//            2048 bytes per k-step), exactly like the A/B tiles of gemm_tcgen05.cu;
//   block B (head_dim 72 only): columns 64..79 as two no-swizzle chunk planes [plane][row/8][row%8][16 B]: plane 0 =
//            columns 64..71 (one small TMA box), plane 1 = zeros (written once), so that the contraction over
//            head_dim stays a multiple of UMMA_K = 16 and N = 16 covers the remaining output columns.
// Every logical MMA becomes "4 k-steps on block A + 1 on block B" (K-major) or "N = 64 on block A + N = 16 on block B"
// per k-step (MN-major).  P is written by the softmax threads in the no-swizzle layout as before.
#include <stdlib.h>
#include <string.h>

#include "attention_tc.cuh"
#include "gemm.h"

namespace mdt {

// operand view of (a 128-row slice of) a split tile
struct SwOp {
  uint32_t a;      // block A address of the slice's first row
  uint32_t b;      // block B plane 0 address of the slice's first row
  uint32_t plane;  // bytes between the two planes of block B (= 16 * rows of the whole tile)
};
MDT_DEVINL SwOp sw_op(uint32_t tile, int tile_rows, int row0) {
  return SwOp{tile + row0 * 128u, tile + tile_rows * 128u + row0 * 16u, tile_rows * 16u};
}
constexpr int sw_tile_bytes(int dp, int rows) { return rows * dp * 2; }

// D[128 x n] = A[128 x DP] * B[n x DP]^T   (both K-major: contraction over head_dim)
template <int DP>
MDT_DEVINL void sw_mma_kk(uint32_t tmem_d, SwOp a, SwOp b, int n) {
  const uint32_t idesc = make_idesc_bf16(kQB, n, 0, 0);
  const uint64_t da = make_smem_desc_sw128(a.a, 16, 1024), db = make_smem_desc_sw128(b.a, 16, 1024);
#pragma unroll
  for (int k = 0; k < 4; ++k) umma_bf16(tmem_d, da + 2 * k, db + 2 * k, idesc, k > 0 ? 1u : 0u);  // +32 B per k-step
  if constexpr (DP > 64)
    umma_bf16(tmem_d, make_smem_desc_nosw(a.b, a.plane, 128), make_smem_desc_nosw(b.b, b.plane, 128), idesc, 1u);
}

// D[128 x DP] (+)= A[128 x 128 tokens] * B[128 tokens x DP]   (B = split tile slice, contraction over its tokens)
// a_desc0 / a_step: descriptor of A's first k-step and its increment (in 16-byte units) per 16 tokens
template <int DP>
MDT_DEVINL void sw_mma_tok(uint32_t tmem_d, uint64_t a_desc0, uint32_t a_step, int a_mn, SwOp b, bool acc0) {
  const uint32_t i64 = make_idesc_bf16(kQB, 64, a_mn, 1), i16 = make_idesc_bf16(kQB, 16, a_mn, 1);
  const uint64_t db = make_smem_desc_sw128(b.a, 8192, 1024);
  const uint64_t db2 = make_smem_desc_nosw(b.b, 128, b.plane);
#pragma unroll
  for (int k = 0; k < kQB / 16; ++k) {
    const uint64_t da = a_desc0 + static_cast<uint64_t>(k) * a_step;
    const uint32_t acc = (acc0 || k > 0) ? 1u : 0u;
    umma_bf16(tmem_d, da, db + static_cast<uint64_t>(k) * (2048 >> 4), i64, acc);
    if constexpr (DP > 64) umma_bf16(tmem_d + 64, da, db2 + static_cast<uint64_t>(k) * (256 >> 4), i16, acc);
  }
}

constexpr int sw_fwd_v_offset(int dp, int tk) {
  const int qk = (kQB + tk) * dp * 2, pb = kQB * tk * 2;
  return qk > pb ? qk : pb;
}
constexpr int sw_fwd_tpr(int tk) { return tk >= 256 ? 2 : 1; }

template <int DP, int TK>
__global__ void __launch_bounds__(kQB * sw_fwd_tpr(TK))
attn_sw_fwd_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                   __nv_bfloat16* __restrict__ out, float* __restrict__ lse, int T, int H, int dh, float scale) {
  constexpr int kPBlk = (TK / 8) * 128;
  constexpr int kTmemCols = TK;  // O aliases S
  constexpr int kVOff = sw_fwd_v_offset(DP, TK);
  constexpr int kTPR = sw_fwd_tpr(TK), kHalf = TK / kTPR;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // [Q | K] (later P), V behind whichever is larger - as in attn_tc_fwd_kernel; all tile bases are multiples of 1024
  const uint32_t sQ = smem_u32(smem), sK = sQ + sw_tile_bytes(DP, kQB), sV = sQ + kVOff, sP = sQ;
  float* s_red = reinterpret_cast<float*>(smem + kVOff + sw_tile_bytes(DP, TK));  // [2][128] max, [2][128] sum
  uint64_t* bar = reinterpret_cast<uint64_t*>(s_red + 4 * kQB);
  uint64_t* ld_bar = bar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 2);

  const int tid = threadIdx.x, warp = tid >> 5, row = tid & (kQB - 1), half = tid >> 7;
  const int b = blockIdx.y / H, h = blockIdx.y % H, q0 = blockIdx.x * kQB;
  if (warp == 0) tmem_alloc<kTmemCols>(tmem_slot);
  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_init(ld_bar, 1);
    fence_barrier_init();
  }
  __syncthreads();  // barrier init visible to the issuing warp
  if (warp == 1) {
    const int lane = tid & 31;
    if (lane == 0) mbar_arrive_expect_tx(ld_bar, static_cast<uint32_t>((kQB + 2 * TK) * dh * 2));
    __syncwarp();
    // lanes 0..(1 + 2 TK/128): block A boxes (64 x 128 elements); the next ones: block B boxes (8 columns x 128 rows)
    constexpr int kBoxes = 1 + 2 * (TK / kQB);
    const int row_g = b * T;
    if (lane < kBoxes || (DP > 64 && lane < 2 * kBoxes)) {
      const int i = lane < kBoxes ? lane : lane - kBoxes;
      const int sel = i == 0 ? 0 : (i <= TK / kQB ? 1 : 2);
      const int blk = i == 0 ? 0 : (sel == 1 ? i - 1 : i - 1 - TK / kQB);  // 128-row block inside the K / V tile
      const uint32_t tile = sel == 0 ? sQ : (sel == 1 ? sK : sV);
      const int tile_rows = sel == 0 ? kQB : TK;
      const int r0 = row_g + (sel == 0 ? q0 : blk * kQB);
      const SwOp op = sw_op(tile, tile_rows, blk * kQB);
      if (lane < kBoxes) tma_load_2d(&tm_a, ld_bar, smem + (op.a - sQ), (sel * H + h) * dh, r0);
      else tma_load_4d(&tm_b, ld_bar, op.b, 0, 0, ((sel * H + h) * dh) / 8 + 8, r0 / 8);
    }
  }
  if constexpr (DP > 64) {  // plane 1 of every block B: zero columns 72..79
    for (int r = tid; r < kQB + 2 * TK; r += blockDim.x) {
      const uint32_t tile = r < kQB ? sQ : (r < kQB + TK ? sK : sV);
      const int tile_rows = r < kQB ? kQB : TK;
      const int rr = r < kQB ? r : (r < kQB + TK ? r - kQB : r - kQB - TK);
      sts128u(tile + tile_rows * 128 + tile_rows * 16 + rr * 16, make_uint4(0, 0, 0, 0));
    }
    fence_proxy_async_smem();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tS = tmem, tO = tmem;
  if (tid == 0) {
    mbar_wait(ld_bar, 0);
    sw_mma_kk<DP>(tS, sw_op(sQ, kQB, 0), sw_op(sK, TK, 0), TK);
    umma_commit(bar);
  }
  mbar_wait(bar, 0);
  tcgen05_fence_after();

  // softmax of this thread's (half) row, straight out of TMEM (lane = row)
  const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) * 32) << 16;
  const float sl = scale * 1.4426950408889634f;
  const int c_lo = half * kHalf;
  float m = -INFINITY;
#pragma unroll 1
  for (int c = c_lo; c < c_lo + kHalf; c += 32) {
    uint32_t r[32];
    tmem_ld_32x32b_x32(tS + lane_addr + c, r);
    tcgen05_wait_ld();
#pragma unroll
    for (int j = 0; j < 32; ++j) m = fmaxf(m, __uint_as_float(r[j]));
  }
  if constexpr (kTPR == 2) {
    s_red[half * kQB + row] = m;
    __syncthreads();
    m = fmaxf(s_red[row], s_red[kQB + row]);
  }
  const float msl = m * sl;
  float l = 0.f;
  const uint32_t prow = sP + (row >> 3) * kPBlk + (row & 7) * 16;
#pragma unroll 1
  for (int c = c_lo; c < c_lo + kHalf; c += 32) {
    uint32_t r[32];
    tmem_ld_32x32b_x32(tS + lane_addr + c, r);
    tcgen05_wait_ld();
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      float p[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        p[j] = fast_exp2(__uint_as_float(r[8 * g + j]) * sl - msl);
        l += p[j];
      }
      sts128u(prow + (c / 8 + g) * 128,
              make_uint4(pack_bf16(p[0], p[1]), pack_bf16(p[2], p[3]), pack_bf16(p[4], p[5]), pack_bf16(p[6], p[7])));
    }
  }
  if constexpr (kTPR == 2) s_red[(2 + half) * kQB + row] = l;
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  if (tid == 0) {
    tcgen05_fence_after();
    // O[128 x DP] = P[128 x TK] (K-major, no-swizzle) * V (split tile, contraction over the keys)
#pragma unroll
    for (int kb = 0; kb < TK / kQB; ++kb)
      sw_mma_tok<DP>(tO, make_smem_desc_nosw(sP + kb * (kQB / 8) * 128, 128, kPBlk), 256 >> 4, 0, sw_op(sV, TK, kb * kQB),
                     kb > 0);
    umma_commit(bar);
  }
  if constexpr (kTPR == 2) l = s_red[2 * kQB + row] + s_red[3 * kQB + row];
  const float inv_l = 1.f / l;
  const int q = q0 + row;
  __nv_bfloat16* orow = out + (static_cast<long long>(b) * T + q) * (H * dh) + h * dh;
  if (lse && half == 0) lse[(static_cast<long long>(b) * H + h) * T + q] = m * scale + logf(l);
  mbar_wait(bar, 1);
  tcgen05_fence_after();
  {
    constexpr int HC = DP / kTPR;
    uint32_t r[HC];
    tmem_ld_cols<HC>(tO + lane_addr + half * HC, r);
    store_row_bf16<HC>(orow, half * HC, r, dh, inv_l);
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) {
    tcgen05_fence_after();
    tmem_dealloc<kTmemCols>(tmem);
  }
}

// ------------------------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------------------------
static bool attn_sw() {  // MDT_ATTN_SW=0: the no-swizzle kernels of attention_tc.cu (A/B switch)
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MDT_ATTN_SW");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

template <int DP, int TK>
static int launch_sw_fwd(const void* qkv, void* out, float* lse, int B, int T, int H, int dh, float scale,
                         cudaStream_t st) {
  const int smem = sw_fwd_v_offset(DP, TK) + sw_tile_bytes(DP, TK) + 4 * kQB * 4 + 64 + 1024;
  auto kern = attn_sw_fwd_kernel<DP, TK>;
  static bool set = false;
  if (!set) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return MDT_ERR_CUDA;
    set = true;
  }
  alignas(64) CUtensorMap ta, tb;
  const unsigned long long rows = static_cast<unsigned long long>(B) * T, cols = 3ull * H * dh;
  int rc = make_row_tile_tmap(&ta, qkv, rows, cols, 64, kQB);
  if (rc != MDT_OK) return rc;
  if (DP > 64) {
    rc = make_token_tile_tmap(&tb, qkv, rows, cols, 1, kQB / 8);
    if (rc != MDT_OK) return rc;
  } else {
    memcpy(&tb, &ta, sizeof(tb));
  }
  kern<<<dim3(T / kQB, B * H), kQB * sw_fwd_tpr(TK), smem, st>>>(ta, tb, static_cast<__nv_bfloat16*>(out), lse, T, H,
                                                                  dh, scale);
  return cudaGetLastError() == cudaSuccess ? MDT_OK : MDT_ERR_CUDA;
}

int attention_sw_fwd(const void* qkv, void* out, float* lse, int B, int T, int H, int dh, float scale,
                     cudaStream_t st) {
  if (!attn_sw() || (reinterpret_cast<uintptr_t>(qkv) & 15)) return MDT_ERR_UNSUPPORTED;
  if (dh == 72) {
    if (T == 128) return launch_sw_fwd<80, 128>(qkv, out, lse, B, T, H, dh, scale, st);
    if (T == 256) return launch_sw_fwd<80, 256>(qkv, out, lse, B, T, H, dh, scale, st);
  } else if (dh == 64) {
    if (T == 128) return launch_sw_fwd<64, 128>(qkv, out, lse, B, T, H, dh, scale, st);
    if (T == 256) return launch_sw_fwd<64, 256>(qkv, out, lse, B, T, H, dh, scale, st);
  }
  return MDT_ERR_UNSUPPORTED;
}

}  // namespace mdt
