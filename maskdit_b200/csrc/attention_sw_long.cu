// Blocked attention (T = NB * 128: the 512-px configs run the encoder at T = 512, head_dim 72, and the decoder at
// T = 1024, head_dim 32) on the split TMA tiles of attention_sw.cuh.  Same algorithms as attention_tc_long.cu -
// two-pass forward with K / V resident, dQ kernel streaming K / V blocks, dK/dV kernel streaming Q / dO blocks, delta
// from the shared pre-kernel - with the tile fills done by bulk tensor copies (one 128-row box per block and tile part)
// and the gradient tiles staged in smem and written by bulk tensor stores.
//
// STATUS: default path for T = 512 / 1024 (and the T = 256, head_dim 64/72 backward) since round 2: parity green on
// B200 (tests/test_kernels_gpu.py::test_attention_fwd_bwd asserts this file's kernels ran), measured at B=128:
// T=512,d_h=72 fwd 721 us / bwd 1601 us (attention_tc_long.cu: 860 / 1963); T=1024,d_h=32 fwd 1776 / bwd 3617 us
// (2173 / 4311).  MDT_ATTN_SWL=0 switches back to attention_tc_long.cu for A/B runs.
#include <stdlib.h>
#include <string.h>

#include "attention_sw.cuh"
#include "gemm.h"

namespace mdt {

void launch_attn_delta(const void* out, const void* dout, float* delta, int B, int T, int H, int dh, cudaStream_t st);

constexpr int kSwlThreads = 256;  // two threads per query / key row (row = tid & 127, column half = tid >> 7)
constexpr float kSwlLog2e = 1.4426950408889634f;

// one 128-row block of a split tile: block A box (+ block B box for head_dim 72); issued by the calling thread
template <int DP>
MDT_DEVINL void swl_load_block(const CUtensorMap* ta, const CUtensorMap* tb, uint64_t* bar, uint8_t* smem, uint32_t s0,
                               uint32_t tile, int tile_rows, int row0, int col, int grow) {
  const SwOp op = sw_op(tile, tile_rows, row0, sw_row_bytes(DP));
  tma_load_2d(ta, bar, smem + (op.a - s0), col, grow);
  if constexpr (DP > 64) tma_load_4d(tb, bar, op.b, 0, 0, col / 8 + 8, grow / 8);
}
template <int DP>
MDT_DEVINL void swl_store_block(const CUtensorMap* ta, const CUtensorMap* tb, uint32_t tile, int col, int grow) {
  tma_store_2d(ta, tile, col, grow);
  if constexpr (DP > 64) tma_store_4d(tb, tile + kQB * 128, 0, 0, col / 8 + 8, grow / 8);
}
// zero plane 1 (columns 72..79) of `rows` rows of a tile with `tile_rows` rows (head_dim 72 only)
template <int DP>
MDT_DEVINL void swl_zero_pad(uint32_t tile, int tile_rows, int rows) {
  if constexpr (DP > 64)
    for (int r = threadIdx.x; r < rows; r += blockDim.x)
      sts128u(tile + tile_rows * 128 + tile_rows * 16 + r * 16, make_uint4(0, 0, 0, 0));
}
// N fp32 values of one row -> bf16 chunks of a 128-row gradient tile in the layout of its head_dim
template <int DP, int N>
MDT_DEVINL void swl_stage_row(uint32_t tile, int row, int c8_0, const uint32_t* r, int dh) {
  if constexpr (DP >= 64) stage_row_split<DP, N>(tile, row, c8_0, r, dh);
  else stage_row_sw64<N>(tile, row, c8_0, r);
}

// ------------------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------------------
template <int DP, int NB>
__global__ void __launch_bounds__(kSwlThreads, 1)
attn_swl_fwd_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                    __nv_bfloat16* __restrict__ out, float* __restrict__ lse, int H, int dh, float scale) {
  constexpr int T = NB * kQB;
  constexpr int kPBlk = (kQB / 8) * 128;
  constexpr uint32_t kRow = sw_row_bytes(DP);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t sQ = smem_u32(smem), sK = sQ + sw_tile_bytes(DP, kQB), sV = sK + sw_tile_bytes(DP, T),
                 sP = sV + sw_tile_bytes(DP, T);
  float* s_red = reinterpret_cast<float*>(smem + sw_tile_bytes(DP, kQB) + 2 * sw_tile_bytes(DP, T) + kQB * kQB * 2);
  uint64_t* bar = reinterpret_cast<uint64_t*>(s_red + 2 * kQB);
  uint64_t* ld_bar = bar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, row = tid & (kQB - 1), half = tid >> 7;
  const int b = blockIdx.y / H, h = blockIdx.y % H, q0 = blockIdx.x * kQB;
  if (warp == 0) tmem_alloc<512>(tmem_slot);
  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_init(ld_bar, 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (warp == 1) {
    if (lane == 0) mbar_arrive_expect_tx(ld_bar, static_cast<uint32_t>((kQB + 2 * T) * dh * 2));
    __syncwarp();
    if (lane < 1 + 2 * NB) {  // lane 0: Q block; lanes 1..NB: K blocks; lanes NB+1..2NB: V blocks
      const int sel = lane == 0 ? 0 : (lane <= NB ? 1 : 2);
      const int blk = lane == 0 ? 0 : (sel == 1 ? lane - 1 : lane - 1 - NB);
      const uint32_t tile = sel == 0 ? sQ : (sel == 1 ? sK : sV);
      swl_load_block<DP>(&tm_a, &tm_b, ld_bar, smem, sQ, tile, sel == 0 ? kQB : T, blk * kQB, (sel * H + h) * dh,
                         b * T + (sel == 0 ? q0 : blk * kQB));
    }
  }
  swl_zero_pad<DP>(sQ, kQB, kQB);
  swl_zero_pad<DP>(sK, T, T);
  swl_zero_pad<DP>(sV, T, T);
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tS = tmem, tO = tmem + 256;  // S blocks ping-pong between columns 0..127 and 128..255
  const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) * 32) << 16;
  const float sl = scale * kSwlLog2e;
  const SwOp oQ = sw_op(sQ, kQB, 0, kRow);
  uint32_t phase = 0;

  // ---- pass 1: row maximum over all key blocks ----
  if (tid == 0) {
    mbar_wait(ld_bar, 0);
    sw_mma_kk<DP>(tS, oQ, sw_op(sK, T, 0, kRow), kQB);
    umma_commit(bar);
  }
  float m = -INFINITY;
  for (int j = 0; j < NB; ++j) {
    mbar_wait(bar, phase);
    phase ^= 1;
    tcgen05_fence_after();
    if (tid == 0 && j + 1 < NB) {  // next block into the other S buffer (its readers finished before the last sync)
      sw_mma_kk<DP>(tS + ((j + 1) & 1) * 128, oQ, sw_op(sK, T, (j + 1) * kQB, kRow), kQB);
      umma_commit(bar);
    }
#pragma unroll 1
    for (int c = half * 64; c < half * 64 + 64; c += 32) {
      uint32_t r[32];
      tmem_ld_32x32b_x32(tS + (j & 1) * 128 + lane_addr + c, r);
      tcgen05_wait_ld();
#pragma unroll
      for (int i = 0; i < 32; ++i) m = fmaxf(m, __uint_as_float(r[i]));
    }
    tcgen05_fence_before();
    __syncthreads();
  }
  s_red[half * kQB + row] = m;
  __syncthreads();
  m = fmaxf(s_red[row], s_red[kQB + row]);
  const float msl = m * sl;
  __syncthreads();  // s_red is reused for the row sums below

  // ---- pass 2: P = exp(S - max), O += P V ----
  if (tid == 0) {
    tcgen05_fence_after();
    sw_mma_kk<DP>(tS, oQ, sw_op(sK, T, 0, kRow), kQB);
    umma_commit(bar);
  }
  float l = 0.f;
  const uint32_t prow = sP + (row >> 3) * kPBlk + (row & 7) * 16;
  for (int j = 0; j < NB; ++j) {
    mbar_wait(bar, phase);  // S_j ready, and P V_{j-1} (which read sP) complete
    phase ^= 1;
    tcgen05_fence_after();
#pragma unroll 1
    for (int c = half * 64; c < half * 64 + 64; c += 32) {
      uint32_t r[32];
      tmem_ld_32x32b_x32(tS + (j & 1) * 128 + lane_addr + c, r);
      tcgen05_wait_ld();
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        float p[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          p[i] = fast_exp2(__uint_as_float(r[8 * g + i]) * sl - msl);
          l += p[i];
        }
        sts128u(prow + (c / 8 + g) * 128,
                make_uint4(pack_bf16(p[0], p[1]), pack_bf16(p[2], p[3]), pack_bf16(p[4], p[5]), pack_bf16(p[6], p[7])));
      }
    }
    fence_proxy_async_smem();
    tcgen05_fence_before();
    __syncthreads();
    if (tid == 0) {
      tcgen05_fence_after();
      sw_mma_tok<DP>(tO, make_smem_desc_nosw(sP, 128, kPBlk), 256 >> 4, 0, sw_op(sV, T, j * kQB, kRow), j > 0);
      if (j + 1 < NB) sw_mma_kk<DP>(tS + ((j + 1) & 1) * 128, oQ, sw_op(sK, T, (j + 1) * kQB, kRow), kQB);
      umma_commit(bar);
    }
  }
  mbar_wait(bar, phase);
  tcgen05_fence_after();
  s_red[half * kQB + row] = l;
  __syncthreads();
  l = s_red[row] + s_red[kQB + row];
  const float inv_l = 1.f / l;
  const int q = q0 + row;
  __nv_bfloat16* orow = out + (static_cast<long long>(b) * T + q) * (H * dh) + h * dh;
  {
    constexpr int HC = DP / 2;
    uint32_t r[HC];
    tmem_ld_cols<HC>(tO + lane_addr + half * HC, r);
    store_row_bf16<HC>(orow, half * HC, r, dh, inv_l);
  }
  if (lse && half == 0) lse[(static_cast<long long>(b) * H + h) * T + q] = m * scale + logf(l);
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) {
    tcgen05_fence_after();
    tmem_dealloc<512>(tmem);
  }
}

// P / dS of this thread's half row from the S / dP accumulators; writes sdS (and sP if WRITE_P)
template <bool WRITE_P>
MDT_DEVINL void swl_softmax_bwd_half(uint32_t tS, uint32_t tdP, uint32_t lane_addr, int half, int row, uint32_t sP,
                                     uint32_t sdS, float sl, float lsl, float delta, float scale) {
  constexpr int kPBlk = (kQB / 8) * 128;
  const uint32_t prow = (row >> 3) * kPBlk + (row & 7) * 16;
#pragma unroll 1
  for (int c = half * 64; c < half * 64 + 64; c += 32) {
    uint32_t rs_[32], rp[32];
    tmem_ld_32x32b_x32(tS + lane_addr + c, rs_);
    tmem_ld_32x32b_x32(tdP + lane_addr + c, rp);
    tcgen05_wait_ld();
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      float p[8], ds[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        p[i] = fast_exp2(__uint_as_float(rs_[8 * g + i]) * sl - lsl);
        ds[i] = p[i] * (__uint_as_float(rp[8 * g + i]) - delta) * scale;
      }
      const uint32_t o = prow + (c / 8 + g) * 128;
      if (WRITE_P)
        sts128u(sP + o, make_uint4(pack_bf16(p[0], p[1]), pack_bf16(p[2], p[3]), pack_bf16(p[4], p[5]),
                                   pack_bf16(p[6], p[7])));
      sts128u(sdS + o, make_uint4(pack_bf16(ds[0], ds[1]), pack_bf16(ds[2], ds[3]), pack_bf16(ds[4], ds[5]),
                                  pack_bf16(ds[6], ds[7])));
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// backward, dQ: CTA = (query block, (b,h)); K / V blocks streamed through a two-stage ring
// ------------------------------------------------------------------------------------------------------------
template <int DP, int NB>
__global__ void __launch_bounds__(kSwlThreads, 1)
attn_swl_dq_kernel(const __grid_constant__ CUtensorMap tq_a, const __grid_constant__ CUtensorMap tq_b,
                   const __grid_constant__ CUtensorMap td_a, const __grid_constant__ CUtensorMap td_b,
                   const __grid_constant__ CUtensorMap tg_a, const __grid_constant__ CUtensorMap tg_b,
                   const float* __restrict__ lse, const float* __restrict__ delta_g, int H, int dh, float scale) {
  constexpr int T = NB * kQB;
  constexpr int kPBlk = (kQB / 8) * 128;
  constexpr uint32_t kRow = sw_row_bytes(DP);
  constexpr int kBlk = sw_tile_bytes(DP, kQB);  // one 128-row tile
  static_assert(kBlk <= kQB * kQB * 2, "dQ is staged in the dS region");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t s0 = smem_u32(smem);
  const uint32_t sQ = s0, sdO = s0 + kBlk, sKV = s0 + 2 * kBlk;  // sKV: 2 x (K block | V block)
  const uint32_t sdS = sKV + 4 * kBlk;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 6 * kBlk + kQB * kQB * 2);
  uint64_t* ld_bar = bar + 1;  // [2]: one per ring stage (stage 0 also carries Q and dO)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 3);

  const int tid = threadIdx.x, warp = tid >> 5, row = tid & (kQB - 1), half = tid >> 7;
  const int b = blockIdx.y / H, h = blockIdx.y % H, qb = blockIdx.x;
  if (warp == 0) tmem_alloc<512>(tmem_slot);
  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_init(&ld_bar[0], 1);
    mbar_init(&ld_bar[1], 1);
    fence_barrier_init();
  }
  for (int t = 0; t < 6; ++t) swl_zero_pad<DP>(s0 + t * kBlk, kQB, kQB);
  fence_proxy_async_smem();
  __syncthreads();
  const int grow = b * T;
  auto load_kv = [&](int kb) {  // tid 0 only
    const uint32_t dst = sKV + (kb & 1) * 2 * kBlk;
    swl_load_block<DP>(&tq_a, &tq_b, &ld_bar[kb & 1], smem, s0, dst, kQB, 0, (H + h) * dh, grow + kb * kQB);
    swl_load_block<DP>(&tq_a, &tq_b, &ld_bar[kb & 1], smem, s0, dst + kBlk, kQB, 0, (2 * H + h) * dh, grow + kb * kQB);
  };
  if (tid == 0) {
    mbar_arrive_expect_tx(&ld_bar[0], static_cast<uint32_t>(4 * kQB * dh * 2));
    swl_load_block<DP>(&tq_a, &tq_b, &ld_bar[0], smem, s0, sQ, kQB, 0, h * dh, grow + qb * kQB);
    swl_load_block<DP>(&td_a, &td_b, &ld_bar[0], smem, s0, sdO, kQB, 0, h * dh, grow + qb * kQB);
    load_kv(0);
    if (NB > 1) {
      mbar_arrive_expect_tx(&ld_bar[1], static_cast<uint32_t>(2 * kQB * dh * 2));
      load_kv(1);
    }
  }
  const int q = qb * kQB + row;
  const float lsl = lse[(static_cast<long long>(b) * H + h) * T + q] * kSwlLog2e;
  const float delta = delta_g[(static_cast<long long>(b) * H + h) * T + q];
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tS = tmem, tdP = tmem + 128, tdQ = tmem + 256;
  const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) * 32) << 16;
  const float sl = scale * kSwlLog2e;
  const SwOp oQ = sw_op(sQ, kQB, 0, kRow), odO = sw_op(sdO, kQB, 0, kRow);
  uint32_t phase = 0;

  for (int kb = 0; kb < NB; ++kb) {
    const uint32_t sK = sKV + (kb & 1) * 2 * kBlk, sV = sK + kBlk;
    const SwOp oK = sw_op(sK, kQB, 0, kRow), oV = sw_op(sV, kQB, 0, kRow);
    if (tid == 0) {
      mbar_wait(&ld_bar[kb & 1], (kb >> 1) & 1);
      tcgen05_fence_after();
      sw_mma_kk<DP>(tS, oQ, oK, kQB);
      sw_mma_kk<DP>(tdP, odO, oV, kQB);
      umma_commit(bar);
    }
    mbar_wait(bar, phase);
    phase ^= 1;
    tcgen05_fence_after();
    swl_softmax_bwd_half<false>(tS, tdP, lane_addr, half, row, 0, sdS, sl, lsl, delta, scale);
    fence_proxy_async_smem();
    tcgen05_fence_before();
    __syncthreads();
    if (tid == 0) {
      tcgen05_fence_after();
      sw_mma_tok<DP>(tdQ, make_smem_desc_nosw(sdS, 128, kPBlk), 256 >> 4, 0, oK, kb > 0);  // dQ += dS K[kb]
      umma_commit(bar);
    }
    mbar_wait(bar, phase);  // dS tile and this K/V stage are free again
    phase ^= 1;
    tcgen05_fence_after();
    if (tid == 0 && kb + 2 < NB) {
      mbar_arrive_expect_tx(&ld_bar[kb & 1], static_cast<uint32_t>(2 * kQB * dh * 2));
      load_kv(kb + 2);
    }
  }
  {  // stage dQ as a gradient tile in the dS region, then one bulk store
    constexpr int HC = DP / 2;
    uint32_t r[HC];
    tmem_ld_cols<HC>(tdQ + lane_addr + half * HC, r);
    swl_stage_row<DP, HC>(sdS, row, half * HC / 8, r, dh);
  }
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  if (tid == 0) {
    swl_store_block<DP>(&tg_a, &tg_b, sdS, h * dh, grow + qb * kQB);
    bulk_commit_group();
    bulk_wait_all();
  }
  if (warp == 0) {
    tcgen05_fence_after();
    tmem_dealloc<512>(tmem);
  }
}

// ------------------------------------------------------------------------------------------------------------
// backward, dK / dV: CTA = (key block, (b,h)); Q / dO blocks streamed through a two-stage ring
// ------------------------------------------------------------------------------------------------------------
template <int DP, int NB>
__global__ void __launch_bounds__(kSwlThreads, 1)
attn_swl_dkv_kernel(const __grid_constant__ CUtensorMap tq_a, const __grid_constant__ CUtensorMap tq_b,
                    const __grid_constant__ CUtensorMap td_a, const __grid_constant__ CUtensorMap td_b,
                    const __grid_constant__ CUtensorMap tg_a, const __grid_constant__ CUtensorMap tg_b,
                    const float* __restrict__ lse, const float* __restrict__ delta_g, int H, int dh, float scale) {
  constexpr int T = NB * kQB;
  constexpr int kPBlk = (kQB / 8) * 128;
  constexpr uint32_t kRow = sw_row_bytes(DP);
  constexpr int kBlk = sw_tile_bytes(DP, kQB);
  static_assert(2 * kBlk <= 2 * kQB * kQB * 2, "dK / dV are staged in the P / dS region");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t s0 = smem_u32(smem);
  const uint32_t sK = s0, sV = s0 + kBlk, sQO = s0 + 2 * kBlk;  // sQO: 2 x (Q block | dO block)
  const uint32_t sP = sQO + 4 * kBlk, sdS = sP + kQB * kQB * 2;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 6 * kBlk + 2 * kQB * kQB * 2);
  uint64_t* ld_bar = bar + 1;  // [2]: stage 0 also carries K and V
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 3);

  const int tid = threadIdx.x, warp = tid >> 5, row = tid & (kQB - 1), half = tid >> 7;
  const int b = blockIdx.y / H, h = blockIdx.y % H, kb = blockIdx.x;
  if (warp == 0) tmem_alloc<512>(tmem_slot);
  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_init(&ld_bar[0], 1);
    mbar_init(&ld_bar[1], 1);
    fence_barrier_init();
  }
  for (int t = 0; t < 6; ++t) swl_zero_pad<DP>(s0 + t * kBlk, kQB, kQB);
  fence_proxy_async_smem();
  __syncthreads();
  const int grow = b * T;
  auto load_qo = [&](int qb) {  // tid 0 only
    const uint32_t dst = sQO + (qb & 1) * 2 * kBlk;
    swl_load_block<DP>(&tq_a, &tq_b, &ld_bar[qb & 1], smem, s0, dst, kQB, 0, h * dh, grow + qb * kQB);
    swl_load_block<DP>(&td_a, &td_b, &ld_bar[qb & 1], smem, s0, dst + kBlk, kQB, 0, h * dh, grow + qb * kQB);
  };
  if (tid == 0) {
    mbar_arrive_expect_tx(&ld_bar[0], static_cast<uint32_t>(4 * kQB * dh * 2));
    swl_load_block<DP>(&tq_a, &tq_b, &ld_bar[0], smem, s0, sK, kQB, 0, (H + h) * dh, grow + kb * kQB);
    swl_load_block<DP>(&tq_a, &tq_b, &ld_bar[0], smem, s0, sV, kQB, 0, (2 * H + h) * dh, grow + kb * kQB);
    load_qo(0);
    if (NB > 1) {
      mbar_arrive_expect_tx(&ld_bar[1], static_cast<uint32_t>(2 * kQB * dh * 2));
      load_qo(1);
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tS = tmem, tdP = tmem + 128, tdK = tmem + 256, tdV = tmem + 256 + DP;
  const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) * 32) << 16;
  const float sl = scale * kSwlLog2e;
  const SwOp oK = sw_op(sK, kQB, 0, kRow), oV = sw_op(sV, kQB, 0, kRow);
  uint32_t phase = 0;

  for (int qb = 0; qb < NB; ++qb) {
    const int q = qb * kQB + row;
    const float lsl = lse[(static_cast<long long>(b) * H + h) * T + q] * kSwlLog2e;
    const float delta = delta_g[(static_cast<long long>(b) * H + h) * T + q];
    const uint32_t sQ = sQO + (qb & 1) * 2 * kBlk, sdO = sQ + kBlk;
    const SwOp oQ = sw_op(sQ, kQB, 0, kRow), odO = sw_op(sdO, kQB, 0, kRow);
    if (tid == 0) {
      mbar_wait(&ld_bar[qb & 1], (qb >> 1) & 1);
      tcgen05_fence_after();
      sw_mma_kk<DP>(tS, oQ, oK, kQB);    // S[query, key]
      sw_mma_kk<DP>(tdP, odO, oV, kQB);  // dP[query, key]
      umma_commit(bar);
    }
    mbar_wait(bar, phase);
    phase ^= 1;
    tcgen05_fence_after();
    swl_softmax_bwd_half<true>(tS, tdP, lane_addr, half, row, sP, sdS, sl, lsl, delta, scale);
    fence_proxy_async_smem();
    tcgen05_fence_before();
    __syncthreads();
    if (tid == 0) {
      tcgen05_fence_after();
      // dV += P^T dO[qb] ; dK += dS^T Q[qb]   (contraction over the 128 queries of this block)
      sw_mma_tok<DP>(tdV, make_smem_desc_nosw(sP, kPBlk, 128), (2 * kPBlk) >> 4, 1, odO, qb > 0);
      sw_mma_tok<DP>(tdK, make_smem_desc_nosw(sdS, kPBlk, 128), (2 * kPBlk) >> 4, 1, oQ, qb > 0);
      umma_commit(bar);
    }
    mbar_wait(bar, phase);  // P / dS tiles and this Q/dO stage are free again
    phase ^= 1;
    tcgen05_fence_after();
    if (tid == 0 && qb + 2 < NB) {
      mbar_arrive_expect_tx(&ld_bar[qb & 1], static_cast<uint32_t>(2 * kQB * dh * 2));
      load_qo(qb + 2);
    }
  }
  {  // stage dK (threads 0-127) | dV (threads 128-255) as gradient tiles in the P / dS region
    uint32_t r[DP];
    tmem_ld_cols<DP>((half ? tdV : tdK) + lane_addr, r);
    swl_stage_row<DP, DP>(sP + half * kBlk, row, 0, r, dh);
  }
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  if (tid == 0) {
    swl_store_block<DP>(&tg_a, &tg_b, sP, (H + h) * dh, grow + kb * kQB);
    swl_store_block<DP>(&tg_a, &tg_b, sP + kBlk, (2 * H + h) * dh, grow + kb * kQB);
    bulk_commit_group();
    bulk_wait_all();
  }
  if (warp == 0) {
    tcgen05_fence_after();
    tmem_dealloc<512>(tmem);
  }
}

// ------------------------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------------------------
static bool attn_swl() {  // default since round 2 (B200: T=512 fwd 721 vs 860 us, bwd 1601 vs 1963 us); =0: A/B
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MDT_ATTN_SWL");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

template <int DP>
static int swl_maps(CUtensorMap* ta, CUtensorMap* tb, const void* ptr, unsigned long long rows,
                    unsigned long long cols) {
  int rc = make_row_tile_tmap(ta, ptr, rows, cols, DP >= 64 ? 64 : DP, kQB);
  if (rc != MDT_OK) return rc;
  if (DP > 64) return make_token_tile_tmap(tb, ptr, rows, cols, 1, kQB / 8);
  memcpy(tb, ta, sizeof(CUtensorMap));
  return MDT_OK;
}

template <int DP, int NB>
static int launch_swl_fwd(const void* qkv, void* out, float* lse, int B, int H, int dh, float scale, cudaStream_t st) {
  constexpr int T = NB * kQB;
  constexpr int smem = sw_tile_bytes(DP, kQB) + 2 * sw_tile_bytes(DP, T) + kQB * kQB * 2 + 2 * kQB * 4 + 64 + 1024;
  if constexpr (smem > 232448) {
    return MDT_ERR_UNSUPPORTED;
  } else {
    auto kern = attn_swl_fwd_kernel<DP, NB>;
    static bool set = false;
    if (!set) {
      if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
        return MDT_ERR_CUDA;
      set = true;
    }
    alignas(64) CUtensorMap ta, tb;
    if (int rc = swl_maps<DP>(&ta, &tb, qkv, static_cast<unsigned long long>(B) * T, 3ull * H * dh)) return rc;
    kern<<<dim3(NB, B * H), kSwlThreads, smem, st>>>(ta, tb, static_cast<__nv_bfloat16*>(out), lse, H, dh, scale);
    return cudaGetLastError() == cudaSuccess ? MDT_OK : MDT_ERR_CUDA;
  }
}

template <int DP, int NB>
static int launch_swl_bwd(const void* qkv, const void* out, const void* dout, const float* lse, float* delta,
                          void* dqkv, int B, int H, int dh, float scale, cudaStream_t st) {
  constexpr int T = NB * kQB;
  constexpr int kBlk = sw_tile_bytes(DP, kQB);
  constexpr int smem_dq = 6 * kBlk + kQB * kQB * 2 + 64 + 1024;
  constexpr int smem_dkv = 6 * kBlk + 2 * kQB * kQB * 2 + 64 + 1024;
  auto k_dq = attn_swl_dq_kernel<DP, NB>;
  auto k_dkv = attn_swl_dkv_kernel<DP, NB>;
  static bool set = false;
  if (!set) {
    if (cudaFuncSetAttribute(k_dq, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_dq) != cudaSuccess ||
        cudaFuncSetAttribute(k_dkv, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_dkv) != cudaSuccess)
      return MDT_ERR_CUDA;
    set = true;
  }
  alignas(64) CUtensorMap tm[6];
  const unsigned long long rows = static_cast<unsigned long long>(B) * T;
  if (int rc = swl_maps<DP>(&tm[0], &tm[1], qkv, rows, 3ull * H * dh)) return rc;
  if (int rc = swl_maps<DP>(&tm[2], &tm[3], dout, rows, 1ull * H * dh)) return rc;
  if (int rc = swl_maps<DP>(&tm[4], &tm[5], dqkv, rows, 3ull * H * dh)) return rc;
  launch_attn_delta(out, dout, delta, B, T, H, dh, st);
  k_dq<<<dim3(NB, B * H), kSwlThreads, smem_dq, st>>>(tm[0], tm[1], tm[2], tm[3], tm[4], tm[5], lse, delta, H, dh,
                                                     scale);
  k_dkv<<<dim3(NB, B * H), kSwlThreads, smem_dkv, st>>>(tm[0], tm[1], tm[2], tm[3], tm[4], tm[5], lse, delta, H, dh,
                                                       scale);
  return cudaGetLastError() == cudaSuccess ? MDT_OK : MDT_ERR_CUDA;
}

int attention_sw_long_fwd(const void* qkv, void* out, float* lse, int B, int T, int H, int dh, float scale,
                          cudaStream_t st) {
  if (!attn_swl() || (reinterpret_cast<uintptr_t>(qkv) & 15)) return MDT_ERR_UNSUPPORTED;
  if (T == 512) {
    if (dh == 72) return launch_swl_fwd<80, 4>(qkv, out, lse, B, H, dh, scale, st);
    if (dh == 64) return launch_swl_fwd<64, 4>(qkv, out, lse, B, H, dh, scale, st);
    if (dh == 32) return launch_swl_fwd<32, 4>(qkv, out, lse, B, H, dh, scale, st);
  } else if (T == 1024) {
    if (dh == 32) return launch_swl_fwd<32, 8>(qkv, out, lse, B, H, dh, scale, st);
  }
  return MDT_ERR_UNSUPPORTED;
}

// `delta` = scratch [B,H,T] floats
int attention_sw_long_bwd(const void* qkv, const void* out, const void* dout, const float* lse, float* delta,
                          void* dqkv, int B, int T, int H, int dh, float scale, cudaStream_t st) {
  if (!attn_swl()) return MDT_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(qkv) | reinterpret_cast<uintptr_t>(dout) | reinterpret_cast<uintptr_t>(dqkv)) & 15)
    return MDT_ERR_UNSUPPORTED;
#define MDT_SWLB(DPV, NBV) return launch_swl_bwd<DPV, NBV>(qkv, out, dout, lse, delta, dqkv, B, H, dh, scale, st)
  if (T == 512) {
    if (dh == 72) MDT_SWLB(80, 4);
    if (dh == 64) MDT_SWLB(64, 4);
    if (dh == 32) MDT_SWLB(32, 4);
  } else if (T == 1024) {
    if (dh == 72) MDT_SWLB(80, 8);
    if (dh == 64) MDT_SWLB(64, 8);
    if (dh == 32) MDT_SWLB(32, 8);
  } else if (T == 256) {
    if (dh == 72) MDT_SWLB(80, 2);
    if (dh == 64) MDT_SWLB(64, 2);
  }
#undef MDT_SWLB
  return MDT_ERR_UNSUPPORTED;
}

}  // namespace mdt
