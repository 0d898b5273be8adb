// Attention forward (T = 128 / 256) and backward (T = 128) for head_dim 64 / 72 with TMA-friendly operand tiles
// (reference: timm Attention core, ctor site models/maskdit.py:178; same math as attention_tc.cu).
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

#include "attention_sw.cuh"
#include "gemm.h"

namespace mdt {
extern int g_sm_budget;  // api.cu: SMs the persistent kernels may occupy (0 = all)

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
      const SwOp op = sw_op(tile, tile_rows, blk * kQB, sw_row_bytes(DP));
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
    sw_mma_kk<DP>(tS, sw_op(sQ, kQB, 0, sw_row_bytes(DP)), sw_op(sK, TK, 0, sw_row_bytes(DP)), TK);
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
      sw_mma_tok<DP>(tO, make_smem_desc_nosw(sP + kb * (kQB / 8) * 128, 128, kPBlk), 256 >> 4, 0,
                     sw_op(sV, TK, kb * kQB, sw_row_bytes(DP)), kb > 0);
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
// backward, T = 128 (one key / query block per (b,h)): persistent, one CTA per SM, same phase structure as
// attn_tc_bwd_kernel<DP, 1> (attention_tc.cu) with split tiles: 8 bulk loads per item (Q, K, V, dO: block A + block B),
// gradients staged as split tiles in the dead P / dS region and written by 6 bulk stores.
// ------------------------------------------------------------------------------------------------------------
constexpr int kSwBwdThreads = 256;

template <int DP>
__global__ void __launch_bounds__(kSwBwdThreads, 1)
attn_sw_bwd_kernel(const __grid_constant__ CUtensorMap tm_qkv_a, const __grid_constant__ CUtensorMap tm_qkv_b,
                   const __grid_constant__ CUtensorMap tm_do_a, const __grid_constant__ CUtensorMap tm_do_b,
                   const __grid_constant__ CUtensorMap tm_g_a, const __grid_constant__ CUtensorMap tm_g_b,
                   const float* __restrict__ lse, int H, int dh, float scale, int nitems) {
  constexpr int T = kQB;
  constexpr int kPBlk = (kQB / 8) * 128;
  constexpr int kTileBytes = sw_tile_bytes(DP, T);  // 20 KB (DP = 80) / 16 KB (DP = 64): multiples of 1024
  constexpr int kSetBytes = 4 * kTileBytes;         // Q | K | V | dO
  static_assert(256 + 3 * DP <= 512, "TMEM budget");
  static_assert(3 * kTileBytes <= 2 * kQB * kQB * 2, "gradient tiles are staged in the P / dS region");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t s0 = smem_u32(smem);
  const uint32_t sP = s0 + 2 * kSetBytes, sdS = sP + kQB * kQB * 2;
  float* s_part = reinterpret_cast<float*>(smem + 2 * kSetBytes + 2 * kQB * kQB * 2);  // [2][128] partial deltas
  uint64_t* bar = reinterpret_cast<uint64_t*>(s_part + 2 * kQB);
  uint64_t* ld_bar = bar + 1;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 3);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row = tid & (kQB - 1), half = tid >> 7;
  if (warp == 0) tmem_alloc<512>(tmem_slot);
  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_init(&ld_bar[0], 1);
    mbar_init(&ld_bar[1], 1);
    fence_barrier_init();
  }
  if constexpr (DP > 64) {  // plane 1 (columns 72..79) of every input tile: zero once, TMA never writes it
    for (int r = tid; r < 8 * T; r += blockDim.x)
      sts128u(s0 + (r / T) * kTileBytes + T * 128 + T * 16 + (r % T) * 16, make_uint4(0, 0, 0, 0));
    fence_proxy_async_smem();
  }
  __syncthreads();
  constexpr int kIoWarp = 7;
  auto issue_loads = [&](int item, int set_idx) {
    if (warp != kIoWarp) return;
    const uint32_t set = s0 + set_idx * kSetBytes;
    const int b = item / H, h = item % H;
    if (lane == 0) mbar_arrive_expect_tx(&ld_bar[set_idx], static_cast<uint32_t>(4 * T * dh * 2));
    __syncwarp();
    if (lane < 4 || (DP > 64 && lane < 8)) {
      const int tile = lane & 3;  // 0..2: q, k, v of qkv; 3: dO
      const uint32_t dst = set + tile * kTileBytes;
      const int col = (tile < 3 ? tile * H + h : h) * dh, r0 = b * T;
      if (lane < 4) tma_load_2d(tile < 3 ? &tm_qkv_a : &tm_do_a, &ld_bar[set_idx], smem + (dst - s0), col, r0);
      else tma_load_4d(tile < 3 ? &tm_qkv_b : &tm_do_b, &ld_bar[set_idx], dst + T * 128, 0, 0, col / 8 + 8, r0 / 8);
    }
  };
  int item = blockIdx.x;
  if (item < nitems) issue_loads(item, 0);
  float lse_next = item < nitems ? lse[static_cast<long long>(item) * T + row] : 0.f;
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tS = tmem, tdP = tmem + 128, tdQ = tmem + 256, tdK = tdQ + DP, tdV = tdK + DP;
  const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) * 32) << 16;
  const float sl = scale * 1.4426950408889634f;
  uint32_t phase = 0;

  for (int it = 0; item < nitems; ++it, item += gridDim.x) {
    const uint32_t set = s0 + (it & 1) * kSetBytes;
    const SwOp oQ = sw_op(set, T, 0), oK = sw_op(set + kTileBytes, T, 0), oV = sw_op(set + 2 * kTileBytes, T, 0),
               odO = sw_op(set + 3 * kTileBytes, T, 0);
    const int b = item / H, h = item % H;
    const int nxt = item + gridDim.x;
    const float lsl = lse_next * 1.4426950408889634f;
    // prefetch the next item into the other set: its last readers (the MMAs of item it-1) completed before that
    // item's read-out
    if (nxt < nitems) {
      issue_loads(nxt, (it + 1) & 1);
      lse_next = lse[static_cast<long long>(nxt) * T + row];
    }
    if (tid == 0) {
      mbar_wait(&ld_bar[it & 1], (it >> 1) & 1);
      sw_mma_kk<DP>(tS, oQ, oK, kQB);
      sw_mma_kk<DP>(tdP, odO, oV, kQB);
      umma_commit(bar);
    }
    mbar_wait(bar, phase);
    phase ^= 1;
    tcgen05_fence_after();
    // one pass over S / dP: P stays in registers, delta_q = sum_k P dP meets the other half of the row through smem
    constexpr int kCols = kQB / 2;
    uint32_t pr[kCols], dpr[kCols];
    float part = 0.f;
#pragma unroll
    for (int c = 0; c < kCols; c += 32) {
      tmem_ld_32x32b_x32(tS + lane_addr + half * kCols + c, pr + c);
      tmem_ld_32x32b_x32(tdP + lane_addr + half * kCols + c, dpr + c);
    }
    tcgen05_wait_ld();
#pragma unroll
    for (int j = 0; j < kCols; ++j) {
      const float p = fast_exp2(__uint_as_float(pr[j]) * sl - lsl);
      pr[j] = __float_as_uint(p);
      part = fmaf(p, __uint_as_float(dpr[j]), part);
    }
    s_part[half * kQB + row] = part;
    if (warp == kIoWarp) bulk_wait_read_all();  // previous item's gradient tiles have left the P / dS region
    __syncthreads();
    const float delta = s_part[row] + s_part[kQB + row];
    const uint32_t prow = (row >> 3) * kPBlk + (row & 7) * 16 + half * (kCols / 8) * 128;
#pragma unroll
    for (int g = 0; g < kCols / 8; ++g) {
      float p[8], ds[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        p[j] = __uint_as_float(pr[8 * g + j]);
        ds[j] = p[j] * (__uint_as_float(dpr[8 * g + j]) - delta) * scale;
      }
      sts128u(sP + prow + g * 128, make_uint4(pack_bf16(p[0], p[1]), pack_bf16(p[2], p[3]), pack_bf16(p[4], p[5]),
                                              pack_bf16(p[6], p[7])));
      sts128u(sdS + prow + g * 128, make_uint4(pack_bf16(ds[0], ds[1]), pack_bf16(ds[2], ds[3]),
                                               pack_bf16(ds[4], ds[5]), pack_bf16(ds[6], ds[7])));
    }
    fence_proxy_async_smem();
    tcgen05_fence_before();
    __syncthreads();
    if (tid == 0) {
      tcgen05_fence_after();
      // dV = P^T dO ; dK = dS^T Q (contraction over the queries) ; dQ = dS K (contraction over the keys)
      sw_mma_tok<DP>(tdV, make_smem_desc_nosw(sP, kPBlk, 128), (2 * kPBlk) >> 4, 1, odO, false);
      sw_mma_tok<DP>(tdK, make_smem_desc_nosw(sdS, kPBlk, 128), (2 * kPBlk) >> 4, 1, oQ, false);
      sw_mma_tok<DP>(tdQ, make_smem_desc_nosw(sdS, 128, kPBlk), 256 >> 4, 0, oK, false);
      umma_commit(bar);
    }
    mbar_wait(bar, phase);
    phase ^= 1;
    tcgen05_fence_after();
    {  // stage dQ (column halves) | dK (threads 0-127) | dV (threads 128-255) as split tiles in the P / dS region
      constexpr int HC = DP / 2;
      uint32_t r[DP];
      tmem_ld_cols<HC>(tdQ + lane_addr + half * HC, r);
      stage_row_split<DP, HC>(sP, row, half * HC / 8, r, dh);
      tmem_ld_cols<DP>((half ? tdV : tdK) + lane_addr, r);
      stage_row_split<DP, DP>(sP + (1 + half) * kTileBytes, row, 0, r, dh);
    }
    fence_proxy_async_smem();
    tcgen05_fence_before();
    __syncthreads();  // accumulators and tiles of this item are dead; the next item may overwrite them
    tcgen05_fence_after();
    if (warp == kIoWarp) {
      if (lane < 3 || (DP > 64 && lane < 6)) {
        const int tile = lane % 3;  // dq, dk, dv
        const uint32_t src = sP + tile * kTileBytes;
        const int col = (tile * H + h) * dh, r0 = b * T;
        if (lane < 3) tma_store_2d(&tm_g_a, src, col, r0);
        else tma_store_4d(&tm_g_b, src + T * 128, 0, 0, col / 8 + 8, r0 / 8);
      }
      bulk_commit_group();
    }
  }
  if (warp == kIoWarp) bulk_wait_all();
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) {
    tcgen05_fence_after();
    tmem_dealloc<512>(tmem);
  }
}

// ------------------------------------------------------------------------------------------------------------
// backward, T = 256, head_dim 32 (decoder): two query blocks x two key blocks per (b,h), persistent, one CTA per SM.
// Same phase structure as attn_tc_bwd_kernel<32, 2> (attention_tc.cu) on SWIZZLE_64B tiles: 10 bulk loads per item
// (Q, K, V, dO, O as [256 x 32] tiles), delta = rowsum(dO * O) read back from smem, dQ staged in the (dead) O tile,
// dK / dV staged in the P region, 6 bulk stores per item.
// Default since round 2 (parity + timing on B200: profiles/r02_attention_paths.md); MDT_ATTN_SW64=0 disables.
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kSwBwdThreads, 1)
attn_sw_bwd2_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_do,
                    const __grid_constant__ CUtensorMap tm_o, const __grid_constant__ CUtensorMap tm_g,
                    const float* __restrict__ lse, int H, float scale, int nitems) {
  constexpr int DP = 32, NB = 2, T = NB * kQB;
  constexpr int kPBlk = (kQB / 8) * 128;
  constexpr uint32_t kRow = 64;                  // bytes per tile row
  constexpr int kTileBytes = T * kRow;           // 16 KB
  constexpr int kSetBytes = 5 * kTileBytes;      // Q | K | V | dO | O
  constexpr int kOutTile = kQB * kRow;           // one staged 128-row gradient tile (8 KB)
  static_assert(4 * kOutTile <= kQB * kQB * 2, "dK / dV tiles are staged in the P region");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t s0 = smem_u32(smem);
  const uint32_t sP = s0 + 2 * kSetBytes, sdS = sP + kQB * kQB * 2;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 2 * kSetBytes + 2 * kQB * kQB * 2);
  uint64_t* ld_bar = bar + 1;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 3);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row = tid & (kQB - 1), half = tid >> 7;
  if (warp == 0) tmem_alloc<512>(tmem_slot);
  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_init(&ld_bar[0], 1);
    mbar_init(&ld_bar[1], 1);
    fence_barrier_init();
  }
  __syncthreads();
  constexpr int kIoWarp = 7;
  auto issue_loads = [&](int item, int set_idx) {  // 5 tiles x 2 boxes of 128 rows
    if (warp != kIoWarp) return;
    const uint32_t set = s0 + set_idx * kSetBytes;
    const int b = item / H, h = item % H;
    if (lane == 0) mbar_arrive_expect_tx(&ld_bar[set_idx], static_cast<uint32_t>(kSetBytes));
    __syncwarp();
    if (lane < 10) {
      const int tile = lane >> 1, blk = lane & 1;
      const uint32_t dst = set + tile * kTileBytes + blk * kOutTile;
      const int col = (tile < 3 ? tile * H + h : h) * DP, r0 = b * T + blk * kQB;
      const CUtensorMap* m = tile < 3 ? &tm_qkv : (tile == 3 ? &tm_do : &tm_o);
      tma_load_2d(m, &ld_bar[set_idx], smem + (dst - s0), col, r0);
    }
  };
  int item = blockIdx.x;
  if (item < nitems) issue_loads(item, 0);
  float lse_next[NB];
#pragma unroll
  for (int qb = 0; qb < NB; ++qb)
    lse_next[qb] = item < nitems ? lse[static_cast<long long>(item) * T + qb * kQB + row] : 0.f;
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tS = tmem, tdP = tmem + 128, tdQ = tmem + 256, tKV = tmem + 256 + DP;  // dK[j] | dV[j] at tKV + j*2DP
  const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) * 32) << 16;
  const float sl = scale * 1.4426950408889634f;
  uint32_t phase = 0;

  for (int it = 0; item < nitems; ++it, item += gridDim.x) {
    const uint32_t set = s0 + (it & 1) * kSetBytes;
    const uint32_t sQ = set, sK = set + kTileBytes, sV = set + 2 * kTileBytes, sdO = set + 3 * kTileBytes,
                   sO = set + 4 * kTileBytes;
    const int b = item / H, h = item % H;
    const int nxt = item + gridDim.x;
    float lse_all[NB];
#pragma unroll
    for (int qb = 0; qb < NB; ++qb) lse_all[qb] = lse_next[qb];
    // The other set (tiles of item it-1, incl. the dQ tile staged in its O slot) and the P region are read by the bulk
    // stores of item it-1: wait for those reads before the prefetch / the first P write may overwrite them.
    if (warp == kIoWarp) bulk_wait_read_all();
    if (nxt < nitems) {
      issue_loads(nxt, (it + 1) & 1);
#pragma unroll
      for (int qb = 0; qb < NB; ++qb) lse_next[qb] = lse[static_cast<long long>(nxt) * T + qb * kQB + row];
    }
    mbar_wait(&ld_bar[it & 1], (it >> 1) & 1);
    // delta_q = sum_d O[q,d] dO[q,d]: both tiles carry the same chunk permutation, so the row is read position-wise
    float delta_all[NB];
#pragma unroll
    for (int qb = 0; qb < NB; ++qb) {
      const uint32_t off = (qb * kQB + row) * kRow;
      float acc = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint4 a, d;
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w) : "r"(sO + off + c * 16));
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(d.x), "=r"(d.y), "=r"(d.z), "=r"(d.w) : "r"(sdO + off + c * 16));
        acc += bf16_lo(a.x) * bf16_lo(d.x) + bf16_hi(a.x) * bf16_hi(d.x) + bf16_lo(a.y) * bf16_lo(d.y) +
               bf16_hi(a.y) * bf16_hi(d.y) + bf16_lo(a.z) * bf16_lo(d.z) + bf16_hi(a.z) * bf16_hi(d.z) +
               bf16_lo(a.w) * bf16_lo(d.w) + bf16_hi(a.w) * bf16_hi(d.w);
      }
      delta_all[qb] = acc;
    }
    tcgen05_fence_before();
    __syncthreads();  // every thread read O (its slot becomes the dQ staging tile) and the stores above have drained
    if (tid == 0) {
      tcgen05_fence_after();
      sw_mma_kk<DP>(tS, sw_op(sQ, T, 0, kRow), sw_op(sK, T, 0, kRow), kQB);
      sw_mma_kk<DP>(tdP, sw_op(sdO, T, 0, kRow), sw_op(sV, T, 0, kRow), kQB);
      umma_commit(bar);
    }
#pragma unroll 1
    for (int qb = 0; qb < NB; ++qb) {
      const float delta = qb == 0 ? delta_all[0] : delta_all[NB - 1];
      const float lsl = (qb == 0 ? lse_all[0] : lse_all[NB - 1]) * 1.4426950408889634f;
#pragma unroll 1
      for (int kb = 0; kb < NB; ++kb) {
        mbar_wait(bar, phase);
        phase ^= 1;
        tcgen05_fence_after();
        const uint32_t prow = (row >> 3) * kPBlk + (row & 7) * 16;
#pragma unroll 1
        for (int c = half * (kQB / 2); c < (half + 1) * (kQB / 2); c += 32) {
          uint32_t rs_[32], rp[32];
          tmem_ld_32x32b_x32(tS + lane_addr + c, rs_);
          tmem_ld_32x32b_x32(tdP + lane_addr + c, rp);
          tcgen05_wait_ld();
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            float p[8], ds[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              p[j] = fast_exp2(__uint_as_float(rs_[8 * g + j]) * sl - lsl);
              ds[j] = p[j] * (__uint_as_float(rp[8 * g + j]) - delta) * scale;
            }
            const uint32_t o = prow + (c / 8 + g) * 128;
            sts128u(sP + o, make_uint4(pack_bf16(p[0], p[1]), pack_bf16(p[2], p[3]), pack_bf16(p[4], p[5]),
                                       pack_bf16(p[6], p[7])));
            sts128u(sdS + o, make_uint4(pack_bf16(ds[0], ds[1]), pack_bf16(ds[2], ds[3]), pack_bf16(ds[4], ds[5]),
                                        pack_bf16(ds[6], ds[7])));
          }
        }
        fence_proxy_async_smem();
        tcgen05_fence_before();
        __syncthreads();
        if (tid == 0) {
          tcgen05_fence_after();
          const uint32_t tdK = tKV + kb * 2 * DP, tdV = tdK + DP;
          // dV[kb] += P^T dO[qb] ; dK[kb] += dS^T Q[qb] (contraction over the queries) ; dQ[qb] += dS K[kb] (over the keys)
          sw_mma_tok<DP>(tdV, make_smem_desc_nosw(sP, kPBlk, 128), (2 * kPBlk) >> 4, 1, sw_op(sdO, T, qb * kQB, kRow), qb > 0);
          sw_mma_tok<DP>(tdK, make_smem_desc_nosw(sdS, kPBlk, 128), (2 * kPBlk) >> 4, 1, sw_op(sQ, T, qb * kQB, kRow), qb > 0);
          sw_mma_tok<DP>(tdQ, make_smem_desc_nosw(sdS, 128, kPBlk), 256 >> 4, 0, sw_op(sK, T, kb * kQB, kRow), kb > 0);
          int nq = qb, nk = kb + 1;
          if (nk == NB) nk = 0, ++nq;
          if (nq < NB) {  // next (qb, kb): S and dP right away, one commit covers everything issued so far
            sw_mma_kk<DP>(tS, sw_op(sQ, T, nq * kQB, kRow), sw_op(sK, T, nk * kQB, kRow), kQB);
            sw_mma_kk<DP>(tdP, sw_op(sdO, T, nq * kQB, kRow), sw_op(sV, T, nk * kQB, kRow), kQB);
          }
          umma_commit(bar);
        }
      }
      // dQ of this query block is complete once the last commit lands (the same commit also covers the next S/dP)
      mbar_wait(bar, phase);
      tcgen05_fence_after();
      {
        constexpr int HC = DP / 2;  // the two threads of a row take 16 columns each = chunks {0,1} / {2,3}
        uint32_t r[HC];
        tmem_ld_cols<HC>(tdQ + lane_addr + half * HC, r);
        stage_row_sw64<HC>(sO, qb * kQB + row, half * (HC / 8), r);
      }
      // the next iteration's first wait uses the same (already completed) phase: do not flip here
      tcgen05_fence_before();
      __syncthreads();  // all rows read dQ before the next query block's MMAs (queued behind this commit) reuse it
    }
    phase ^= 1;  // the last commit of the item has been consumed by the wait above
    // dK / dV: rows = keys; threads 0-127 stage dK, threads 128-255 stage dV: tile (kb, half) in the P region
#pragma unroll 1
    for (int kb = 0; kb < NB; ++kb) {
      uint32_t r[DP];
      tmem_ld_cols<DP>(tKV + kb * 2 * DP + half * DP + lane_addr, r);
      stage_row_sw64<DP>(sP + (kb * 2 + half) * kOutTile, row, 0, r);
    }
    fence_proxy_async_smem();
    tcgen05_fence_before();
    __syncthreads();  // accumulators and tiles of this item are dead; the next item may overwrite them
    tcgen05_fence_after();
    if (warp == kIoWarp) {
      if (lane < 6) {  // dQ blocks 0/1 from the O slot, then (dK, dV) of key blocks 0/1 from the P region
        const int sel = lane < 2 ? 0 : 1 + ((lane - 2) & 1), blk = lane < 2 ? lane : (lane - 2) >> 1;
        const uint32_t src = lane < 2 ? sO + blk * kOutTile : sP + (blk * 2 + (sel - 1)) * kOutTile;
        tma_store_2d(&tm_g, src, (sel * H + h) * DP, b * T + blk * kQB);
      }
      bulk_commit_group();
    }
  }
  if (warp == kIoWarp) bulk_wait_all();
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) {
    tcgen05_fence_after();
    tmem_dealloc<512>(tmem);
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
  int rc = make_row_tile_tmap(&ta, qkv, rows, cols, DP >= 64 ? 64 : DP, kQB);
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
  } else if (dh == 32 && T == 256) {
    // SWIZZLE_64B variant (decoder).  Round 2, B200: parity green, 165 us vs 200 us for the 16-byte tile fills of
    // attention_tc.cu at B=256 (profiles/r02_attention_paths.md); MDT_ATTN_SW64=0 switches back (A/B).
    static const bool on = [] { const char* e = getenv("MDT_ATTN_SW64"); return !(e && e[0] == '0'); }();
    if (on) return launch_sw_fwd<32, 256>(qkv, out, lse, B, T, H, dh, scale, st);
  }
  return MDT_ERR_UNSUPPORTED;
}

template <int DP>
static int launch_sw_bwd(const void* qkv, const void* dout, const float* lse, void* dqkv, int B, int H, int dh,
                         float scale, cudaStream_t st) {
  constexpr int T = kQB;
  const int smem = 8 * sw_tile_bytes(DP, T) + 2 * kQB * kQB * 2 + 2 * kQB * 4 + 64 + 1024;
  auto kern = attn_sw_bwd_kernel<DP>;
  static bool set = false;
  static int sms = 0;
  if (!set) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return MDT_ERR_CUDA;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
    set = true;
  }
  alignas(64) CUtensorMap tm[6];
  const unsigned long long rows = static_cast<unsigned long long>(B) * T;
  const void* ptrs[3] = {qkv, dout, dqkv};
  const unsigned long long cols[3] = {3ull * H * dh, 1ull * H * dh, 3ull * H * dh};
  for (int i = 0; i < 3; ++i) {
    int rc = make_row_tile_tmap(&tm[2 * i], ptrs[i], rows, cols[i], 64, kQB);
    if (rc != MDT_OK) return rc;
    if (DP > 64) {
      rc = make_token_tile_tmap(&tm[2 * i + 1], ptrs[i], rows, cols[i], 1, kQB / 8);
      if (rc != MDT_OK) return rc;
    } else {
      memcpy(&tm[2 * i + 1], &tm[2 * i], sizeof(CUtensorMap));
    }
  }
  const int nitems = B * H;
  const int grid_sms = (g_sm_budget > 0 && g_sm_budget < sms) ? g_sm_budget : sms;   // mdt_set_sm_budget
  kern<<<nitems < grid_sms ? nitems : grid_sms, kSwBwdThreads, smem, st>>>(tm[0], tm[1], tm[2], tm[3], tm[4], tm[5], lse, H, dh,
                                                                 scale, nitems);
  return cudaGetLastError() == cudaSuccess ? MDT_OK : MDT_ERR_CUDA;
}

static int launch_sw_bwd2(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, int B,
                          int H, float scale, cudaStream_t st) {
  constexpr int T = 2 * kQB, DP = 32;
  const int smem = 2 * 5 * T * 64 + 2 * kQB * kQB * 2 + 64 + 1024;
  static bool set = false;
  static int sms = 0;
  if (!set) {
    if (cudaFuncSetAttribute(attn_sw_bwd2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
      return MDT_ERR_CUDA;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
    set = true;
  }
  alignas(64) CUtensorMap tm[4];
  const unsigned long long rows = static_cast<unsigned long long>(B) * T;
  const void* ptrs[4] = {qkv, dout, out, dqkv};
  const unsigned long long cols[4] = {3ull * H * DP, 1ull * H * DP, 1ull * H * DP, 3ull * H * DP};
  for (int i = 0; i < 4; ++i) {
    const int rc = make_row_tile_tmap(&tm[i], ptrs[i], rows, cols[i], DP, kQB);
    if (rc != MDT_OK) return rc;
  }
  const int nitems = B * H;
  const int grid_sms = (g_sm_budget > 0 && g_sm_budget < sms) ? g_sm_budget : sms;
  attn_sw_bwd2_kernel<<<nitems < grid_sms ? nitems : grid_sms, kSwBwdThreads, smem, st>>>(tm[0], tm[1], tm[2], tm[3], lse, H,
                                                                                scale, nitems);
  return cudaGetLastError() == cudaSuccess ? MDT_OK : MDT_ERR_CUDA;
}

int attention_sw_bwd(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, int B, int T, int H, int dh,
                     float scale, cudaStream_t st) {
  if (!attn_sw()) return MDT_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(qkv) | reinterpret_cast<uintptr_t>(dout) | reinterpret_cast<uintptr_t>(dqkv) |
       reinterpret_cast<uintptr_t>(out)) & 15)
    return MDT_ERR_UNSUPPORTED;
  if (T == 2 * kQB && dh == 32) {  // decoder: SWIZZLE_64B tiles (292 us vs 338 us at B=256); MDT_ATTN_SW64=0 = A/B
    static const bool on = [] { const char* e = getenv("MDT_ATTN_SW64"); return !(e && e[0] == '0'); }();
    return on ? launch_sw_bwd2(qkv, out, dout, lse, dqkv, B, H, scale, st) : MDT_ERR_UNSUPPORTED;
  }
  if (T != kQB) return MDT_ERR_UNSUPPORTED;
  if (dh == 72) return launch_sw_bwd<80>(qkv, dout, lse, dqkv, B, H, dh, scale, st);
  if (dh == 64) return launch_sw_bwd<64>(qkv, dout, lse, dqkv, B, H, dh, scale, st);
  return MDT_ERR_UNSUPPORTED;
}

}  // namespace mdt
