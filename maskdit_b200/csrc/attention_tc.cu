// Attention core on the 5th-gen tensor cores (tcgen05.mma, S / dP / O / dQ / dK / dV accumulators in TMEM) for the
// MaskDiT training shapes (head_dim 32 runs here; head_dim 64 / 72 prefers the split-tile kernels of attention_sw.cu): T = 128 or 256 tokens per sample, head_dim 72 (encoder, zero-padded to 80) or 32 (decoder).
// Replaces softmax(q k^T / sqrt(dh)) v of timm Attention (reference ctor site models/maskdit.py:178) and its backward.
//
//   forward : CTA = (128 queries of one (b,h)); S = Q K^T -> TMEM; one thread per query row does the softmax straight
//             out of TMEM (two passes, fp32), writes P (bf16) to smem; O = P V -> TMEM -> bf16 rows + log-sum-exp.
//   backward: CTA = one (b,h); per (query block, key block): S = Q K^T and dP = dO V^T -> TMEM; row threads form
//             P = exp(S - lse), dS = P (dP - delta) scale -> smem (bf16); dV += P^T dO, dK += dS^T Q, dQ += dS K
//             accumulate in TMEM over the loop.  No atomics, no recompute pass, deterministic.
//
// Shared-memory operand tiles use the UMMA canonical NO-SWIZZLE layout: 8x8 "core matrices" of 128 contiguous bytes
// (8 rows of 16 B).  A token-major tile [rows x DP] is stored as  off(row, c8) = (row/8)*ROWBLK + c8*128 + (row%8)*16
// with ROWBLK = (DP/8)*128.  The SAME bytes serve as a K-major operand (contraction over head_dim: LBO = 128,
// SBO = ROWBLK) and as an MN-major operand (contraction over tokens: SBO = 128, LBO = ROWBLK), so Q, K, V, dO, P and dS
// are each staged exactly once.  head_dim 72 -> 9 real 16-byte chunks + 1 zero chunk per row (144-byte rows cannot
// be TMA-swizzled; the tiles are filled with coalesced 16-byte loads instead).
#include <stdlib.h>
#include <string.h>

#include "attention_tc.cuh"
#include "gemm.h"
#include "../../include/maskdit_b200.h"

namespace mdt {

// ------------------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------------------
constexpr int fwd_v_offset(int dp, int tk) {
  const int qk = (kQB + tk) * dp * 2, pb = kQB * tk * 2;
  return qk > pb ? qk : pb;
}

// threads per query row: T = 256 splits each row's key columns over two threads (row = tid & 127); for T = 128 one
// thread per row was measured faster (114 vs 121 us at B=256, d_h=72: the extra barrier costs more than it saves)
constexpr int fwd_tpr(int tk) { return tk >= 256 ? 2 : 1; }

// Tile fill.  r01 phase timing (clock64 around the phases of a one-CTA-per-SM variant): ISSUING the 16-byte cp.async
// of one item (60 KB at T = 128, d_h = 72) took 3.9k of the 8.9k cycles an item needs - the LSU accepts ~16 B/clk
// per SM of such requests - and a persistent double-buffered variant was slower than this kernel for that reason.
// With kTMA the tiles are written by TMA instead (4-D tensor map, see make_token_tile_tmap: the box lands directly
// in the core-matrix layout): one warp issues <= 3 bulk copies per lane and nothing else touches the LSU.  Rows of
// d_h = 72 are copied as 9 chunks per 8-row block (one box each) so that the zero pad chunk stays zero.
template <int DP, int TK, bool kTMA>
__global__ void __launch_bounds__(kQB * fwd_tpr(TK))
attn_tc_fwd_kernel(const __nv_bfloat16* __restrict__ qkv, const __grid_constant__ CUtensorMap tm_qkv,
                   __nv_bfloat16* __restrict__ out, float* __restrict__ lse, int T, int H, int dh, float scale) {
  using TT = TokTile<DP>;
  constexpr int kPBlk = (TK / 8) * 128;                       // bytes per 8-query block of P
  constexpr int kTmemCols = TK;  // O aliases S: S is dead once every row thread has written its P row to smem
  extern __shared__ __align__(128) uint8_t smem[];
  // P overwrites Q|K: both are dead once the S MMAs have completed (the row threads only start writing P after
  // waiting on that commit); V starts behind whichever of the two is larger.  T = 128: 92 KB -> 60 KB of smem (3 CTAs
  // per SM); T = 256, d_h = 72: 164 KB -> 104 KB (2 CTAs per SM, which is also what the 256 TMEM columns allow).
  constexpr int kVOff = fwd_v_offset(DP, TK);
  const uint32_t sQ = smem_u32(smem), sK = sQ + kQB * DP * 2, sV = sQ + kVOff;
  const uint32_t sP = sQ;
  float* s_red = reinterpret_cast<float*>(smem + kVOff + TK * DP * 2);  // [2][128] row max, [2][128] row sum
  uint64_t* bar = reinterpret_cast<uint64_t*>(s_red + 4 * kQB);
  uint64_t* ld_bar = bar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 2);

  const int tid = threadIdx.x, warp = tid >> 5, row = tid & (kQB - 1), half = tid >> 7;
  const int b = blockIdx.y / H, h = blockIdx.y % H, q0 = blockIdx.x * kQB;
  const long long rs = 3LL * H * dh;
  const __nv_bfloat16* base = qkv + static_cast<long long>(b) * T * rs;
  if (warp == 0) tmem_alloc<kTmemCols>(tmem_slot);
  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_init(ld_bar, 1);
    fence_barrier_init();
  }
  if constexpr (kTMA) {
    __syncthreads();  // barrier init visible to the issuing warp
    if (warp == 1) {
      const int lane = tid & 31;
      if (lane == 0) mbar_arrive_expect_tx(ld_bar, static_cast<uint32_t>((kQB + 2 * TK) * dh * 2));
      __syncwarp();
      // boxes: whole 128-row tiles when d_h == DP, single 8-row blocks (9 of 10 chunks) otherwise
      const int rb_per_box = (dh == DP) ? kQB / 8 : 1;
      const int nq = (kQB / 8) / rb_per_box, nk = (TK / 8) / rb_per_box;
      const int chunks = dh / 8;
      const int rb0 = static_cast<int>((static_cast<long long>(b) * T) / 8);
      for (int i = lane; i < nq + 2 * nk; i += 32) {
        int sel, j;
        uint32_t dst;
        if (i < nq) sel = 0, j = i, dst = sQ;
        else if (i < nq + nk) sel = 1, j = i - nq, dst = sK;
        else sel = 2, j = i - nq - nk, dst = sV;
        const int rblk = j * rb_per_box;
        tma_load_4d(&tm_qkv, ld_bar, dst + rblk * TT::ROWBLK, 0, 0, (sel * H + h) * chunks,
                    rb0 + (sel == 0 ? q0 / 8 : 0) + rblk);
      }
    }
    if (dh < DP) {  // the pad chunk of every row is never written by TMA: zero it (generic proxy -> fence below)
      for (int r = tid; r < kQB + 2 * TK; r += blockDim.x) {
        const uint32_t tile = r < kQB ? sQ : (r < kQB + TK ? sK : sV);
        const int rr = r < kQB ? r : (r < kQB + TK ? r - kQB : r - kQB - TK);
        for (int c8 = dh / 8; c8 < DP / 8; ++c8) sts128u(tile + TT::off(rr, c8), make_uint4(0, 0, 0, 0));
      }
    }
  } else {
    TT::load(sQ, base + q0 * rs + h * dh, rs, kQB, dh);
    TT::load(sK, base + (H + h) * dh, rs, TK, dh);
    TT::load(sV, base + (2 * H + h) * dh, rs, TK, dh);
    cp_async_wait_all();
  }
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tS = tmem, tO = tmem;
  if (tid == 0) {
    if constexpr (kTMA) mbar_wait(ld_bar, 0);
    mma_kk<DP>(tS, sQ, sK, TK, false);
    umma_commit(bar);
  }
  mbar_wait(bar, 0);
  tcgen05_fence_after();

  // softmax of this thread's half row, straight out of TMEM (lane = row)
  const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) * 32) << 16;
  const float sl = scale * 1.4426950408889634f;
  constexpr int kTPR = fwd_tpr(TK), kHalf = TK / kTPR;
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
    // O[128 x DP] = P[128 x TK] (K-major) * V (MN-major: N = head dim, K = keys)
    const uint32_t idesc = make_idesc_bf16(kQB, DP, 0, 1);
#pragma unroll 4
    for (int k = 0; k < TK / 16; ++k)
      umma_bf16(tO, make_smem_desc_nosw(sP + k * 256, 128, kPBlk),
                make_smem_desc_nosw(sV + k * 2 * TT::ROWBLK, TT::ROWBLK, 128), idesc, k > 0 ? 1u : 0u);
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
// backward (one CTA per (b,h); NB = T / 128 query blocks == key blocks)
// ------------------------------------------------------------------------------------------------------------
constexpr int kBwdThreads = 256;  // two threads per query / key row: each owns half of the columns of every phase

// Persistent: one CTA per SM walks the (b,h) items; the Q/K/V/dO tiles of item i+1 are fetched with cp.async into the
// second tile set while item i is being computed (a per-item CTA spent 8k of its 18k cycles waiting for its tiles).
// delta = rowsum(dO * O):  NB == 1 uses the identity  sum_d dO*O = sum_k P*dP  on the S / dP accumulators already in
// TMEM (no O tile at all); NB == 2 stages O as a fifth tile of the set.
// kTMA: the tile sets are filled by TMA (one warp issues the boxes, completion on a per-set mbarrier) instead of
// cp.async by all threads - see the note at attn_tc_fwd_kernel; the zero pad chunks are written once at kernel start.
template <int DP, int NB, bool kTMA>
__global__ void __launch_bounds__(kBwdThreads, 1)
attn_tc_bwd_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ out,
                   const __nv_bfloat16* __restrict__ dout, const __grid_constant__ CUtensorMap tm_qkv,
                   const __grid_constant__ CUtensorMap tm_do, const __grid_constant__ CUtensorMap tm_o,
                   const __grid_constant__ CUtensorMap tm_dqkv, const float* __restrict__ lse,
                   __nv_bfloat16* __restrict__ dqkv, int H, int dh, float scale, int nitems) {
  using TT = TokTile<DP>;
  constexpr int T = NB * kQB;
  constexpr bool kDeltaFromP = NB == 1;
  constexpr int kPBlk = (kQB / 8) * 128;  // P / dS tiles are [128 queries x 128 keys]
  constexpr int kTileBytes = T * DP * 2;
  constexpr int kSetTiles = kDeltaFromP ? 4 : 5;
  constexpr int kSetBytes = kSetTiles * kTileBytes;
  static_assert(256 + DP + 2 * NB * DP <= 512, "TMEM budget");
  static_assert(NB <= 2, "delta/lse selection below assumes at most two query blocks");
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t s0 = smem_u32(smem);
  const uint32_t sP = s0 + 2 * kSetBytes, sdS = sP + kQB * kQB * 2;
  float* s_part = reinterpret_cast<float*>(smem + 2 * kSetBytes + 2 * kQB * kQB * 2);  // [2][128] partial deltas
  uint64_t* bar = reinterpret_cast<uint64_t*>(s_part + 2 * kQB);
  uint64_t* ld_bar = bar + 1;  // [2]: one per tile set (kTMA)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 3);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int row = tid & (kQB - 1), half = tid >> 7;  // TMEM lane (= row) is fixed by warp % 4, `half` picks the columns
  const long long rs = 3LL * H * dh;
  const int HD = H * dh;
  if (warp == 0) tmem_alloc<512>(tmem_slot);
  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_init(&ld_bar[0], 1);
    mbar_init(&ld_bar[1], 1);
    fence_barrier_init();
  }
  if constexpr (kTMA) {
    if (dh < DP) {  // pad chunks of both sets: zero once, TMA never writes them
      for (int r = tid; r < 2 * kSetTiles * T; r += blockDim.x)
        for (int c8 = dh / 8; c8 < DP / 8; ++c8)
          sts128u(s0 + (r / T) * kTileBytes + TT::off(r % T, c8), make_uint4(0, 0, 0, 0));
      fence_proxy_async_smem();
    }
    __syncthreads();  // barrier init (and the zeroed pads) visible before the first bulk copy is issued
  }
  constexpr int kLoadWarp = 7;
  // NB == 1 with TMA: dQ | dK | dV are staged as bf16 tiles in the (dead) P / dS region and leave through bulk tensor
  // stores.  r01 phase timing: the direct read-out (every thread storing 16-byte pieces of its own row: 32 half-written
  // sectors per warp instruction) took 5.8k of the 15.6k cycles of an item.
  constexpr bool kBulkOut = kTMA && NB == 1;
  auto issue_loads = [&](int item, int set_idx) {
    const uint32_t set = s0 + set_idx * kSetBytes;
    const int b = item / H, h = item % H;
    if constexpr (kTMA) {
      if (warp != kLoadWarp) return;
      const int lane = tid & 31;
      if (lane == 0) mbar_arrive_expect_tx(&ld_bar[set_idx], static_cast<uint32_t>(kSetTiles * T * dh * 2));
      __syncwarp();
      const int rb_per_box = (dh == DP) ? kQB / 8 : 1;
      const int per_tile = (T / 8) / rb_per_box, chunks = dh / 8;
      const int rb0 = static_cast<int>((static_cast<long long>(b) * T) / 8);
      for (int i = lane; i < kSetTiles * per_tile; i += 32) {
        const int tile = i / per_tile, rblk = (i - tile * per_tile) * rb_per_box;
        const uint32_t dst = set + tile * kTileBytes + rblk * TT::ROWBLK;
        if (tile < 3) tma_load_4d(&tm_qkv, &ld_bar[set_idx], dst, 0, 0, (tile * H + h) * chunks, rb0 + rblk);
        else tma_load_4d(tile == 3 ? &tm_do : &tm_o, &ld_bar[set_idx], dst, 0, 0, h * chunks, rb0 + rblk);
      }
    } else {
      const __nv_bfloat16* base = qkv + static_cast<long long>(b) * T * rs;
      TT::load(set, base + h * dh, rs, T, dh);
      TT::load(set + kTileBytes, base + (H + h) * dh, rs, T, dh);
      TT::load(set + 2 * kTileBytes, base + (2 * H + h) * dh, rs, T, dh);
      TT::load(set + 3 * kTileBytes, dout + static_cast<long long>(b) * T * HD + h * dh, HD, T, dh);
      if constexpr (!kDeltaFromP)
        TT::load(set + 4 * kTileBytes, out + static_cast<long long>(b) * T * HD + h * dh, HD, T, dh);
      asm volatile("cp.async.commit_group;" ::: "memory");
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
  constexpr uint32_t RB = TT::ROWBLK;
  constexpr uint32_t kBlkBytes = 16 * RB;  // 128 token rows
  uint32_t phase = 0;

#ifdef MDT_ATTN_PROF  // phase cycle counters (tools/attn_phase_prof.py), written to the delta scratch behind lse
  long long pt[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, t_prev = clock64();
#define MDT_PROF(i) { const long long t_now = clock64(); pt[i] += t_now - t_prev; t_prev = t_now; }
#else
#define MDT_PROF(i)
#endif
  for (int it = 0; item < nitems; ++it, item += gridDim.x) {
    const uint32_t set = s0 + (it & 1) * kSetBytes;
    const uint32_t sQ = set, sK = set + kTileBytes, sV = set + 2 * kTileBytes, sdO = set + 3 * kTileBytes;
    const int b = item / H, h = item % H;
    const int nxt = item + gridDim.x;
    float lse_all[NB];
#pragma unroll
    for (int qb = 0; qb < NB; ++qb) lse_all[qb] = lse_next[qb];
    // prefetch the next item into the other set (its last readers, the MMAs of item it-1, completed before that item's
    // read-out) and this row's next lse values
    if (nxt < nitems) {
      issue_loads(nxt, (it + 1) & 1);
#pragma unroll
      for (int qb = 0; qb < NB; ++qb) lse_next[qb] = lse[static_cast<long long>(nxt) * T + qb * kQB + row];
      if constexpr (!kTMA) asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      if constexpr (!kTMA) asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    if constexpr (kTMA) mbar_wait(&ld_bar[it & 1], (it >> 1) & 1);
    fence_proxy_async_smem();
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    MDT_PROF(0)  // prefetch issue + load wait + barrier

    float delta_all[NB];
    if constexpr (!kDeltaFromP) {
      const uint32_t sO = set + 4 * kTileBytes;
#pragma unroll
      for (int qb = 0; qb < NB; ++qb) {
        const int q = qb * kQB + row;
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < DP / 8; ++c) {
          uint4 a, d;
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w) : "r"(sO + TT::off(q, c)));
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(d.x), "=r"(d.y), "=r"(d.z), "=r"(d.w) : "r"(sdO + TT::off(q, c)));
          acc += bf16_lo(a.x) * bf16_lo(d.x) + bf16_hi(a.x) * bf16_hi(d.x) + bf16_lo(a.y) * bf16_lo(d.y) +
                 bf16_hi(a.y) * bf16_hi(d.y) + bf16_lo(a.z) * bf16_lo(d.z) + bf16_hi(a.z) * bf16_hi(d.z) +
                 bf16_lo(a.w) * bf16_lo(d.w) + bf16_hi(a.w) * bf16_hi(d.w);
        }
        delta_all[qb] = acc;
      }
    }
    if (tid == 0) {  // first S / dP of this item
      mma_kk<DP>(tS, sQ, sK, kQB, false);
      mma_kk<DP>(tdP, sdO, sV, kQB, false);
      umma_commit(bar);
    }
    MDT_PROF(1)  // delta from O (NB = 2) + first S / dP issue
    for (int qb = 0; qb < NB; ++qb) {
      const int q = qb * kQB + row;
      float delta = kDeltaFromP ? 0.f : (qb == 0 ? delta_all[0] : delta_all[NB - 1]);
      const float lsl = (qb == 0 ? lse_all[0] : lse_all[NB - 1]) * 1.4426950408889634f;
      for (int kb = 0; kb < NB; ++kb) {
        mbar_wait(bar, phase);
        phase ^= 1;
        tcgen05_fence_after();
        MDT_PROF(2)  // S / dP MMA wait
        if constexpr (kDeltaFromP) {
          // delta_q = sum_k P[q,k] dP[q,k]; the two threads of a row each sum their 64 keys and meet through smem
          float part = 0.f;
#pragma unroll 1
          for (int c = half * (kQB / 2); c < (half + 1) * (kQB / 2); c += 32) {
            uint32_t rs_[32], rp[32];
            tmem_ld_32x32b_x32(tS + lane_addr + c, rs_);
            tmem_ld_32x32b_x32(tdP + lane_addr + c, rp);
            tcgen05_wait_ld();
#pragma unroll
            for (int j = 0; j < 32; ++j)
              part = fmaf(fast_exp2(__uint_as_float(rs_[j]) * sl - lsl), __uint_as_float(rp[j]), part);
          }
          s_part[half * kQB + row] = part;
          if constexpr (kBulkOut) {
            if (warp == kLoadWarp) bulk_wait_read_all();  // previous item's gradient tiles have left the P / dS region
          }
          __syncthreads();
          delta = s_part[row] + s_part[kQB + row];
        }
        MDT_PROF(3)  // delta pass + exchange
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
        MDT_PROF(4)  // P / dS pass + barrier
        if (tid == 0) {
          tcgen05_fence_after();
          const uint32_t tdK = tKV + kb * 2 * DP, tdV = tdK + DP;
          const uint32_t i_tt = make_idesc_bf16(kQB, DP, 1, 1);  // A, B both token-contracted (MN-major)
          const uint32_t i_kt = make_idesc_bf16(kQB, DP, 0, 1);  // A K-major (keys contiguous), B token-contracted
#pragma unroll
          for (int k = 0; k < kQB / 16; ++k) {
            // dV[kb] += P^T dO[qb] ; dK[kb] += dS^T Q[qb]   (contraction over the 128 queries)
            const uint64_t aP = make_smem_desc_nosw(sP + k * 2 * kPBlk, kPBlk, 128);
            const uint64_t aS = make_smem_desc_nosw(sdS + k * 2 * kPBlk, kPBlk, 128);
            const uint64_t bO = make_smem_desc_nosw(sdO + qb * kBlkBytes + k * 2 * RB, RB, 128);
            const uint64_t bQ = make_smem_desc_nosw(sQ + qb * kBlkBytes + k * 2 * RB, RB, 128);
            umma_bf16(tdV, aP, bO, i_tt, (qb > 0 || k > 0) ? 1u : 0u);
            umma_bf16(tdK, aS, bQ, i_tt, (qb > 0 || k > 0) ? 1u : 0u);
            // dQ[qb] += dS K[kb]   (contraction over the 128 keys)
            const uint64_t aS2 = make_smem_desc_nosw(sdS + k * 256, 128, kPBlk);
            const uint64_t bK = make_smem_desc_nosw(sK + kb * kBlkBytes + k * 2 * RB, RB, 128);
            umma_bf16(tdQ, aS2, bK, i_kt, (kb > 0 || k > 0) ? 1u : 0u);
          }
          // next (qb, kb) of this item: S and dP can be issued right away, one commit covers everything issued so far
          int nq = qb, nk = kb + 1;
          if (nk == NB) nk = 0, ++nq;
          if (nq < NB) {
            mma_kk<DP>(tS, sQ + nq * kBlkBytes, sK + nk * kBlkBytes, kQB, false);
            mma_kk<DP>(tdP, sdO + nq * kBlkBytes, sV + nk * kBlkBytes, kQB, false);
          }
          umma_commit(bar);
        }
        MDT_PROF(5)  // dV / dK / dQ (+ next S / dP) issue
      }
      // dQ of this query block is complete once the last commit lands; the same commit also covers the next S/dP
      mbar_wait(bar, phase);
      tcgen05_fence_after();
      MDT_PROF(6)  // gradient MMA wait
      __nv_bfloat16* grow = dqkv + (static_cast<long long>(b) * T + q) * rs + h * dh;
      {
        // the two threads of a row take the two column halves (DP/2 is a multiple of 8 for DP = 32, 64, 80)
        constexpr int HC = DP / 2;
        uint32_t r[HC];
        tmem_ld_cols<HC>(tdQ + lane_addr + half * HC, r);
        if constexpr (kBulkOut) stage_row_bf16<DP, HC>(sP, row, half * HC / 8, r, dh);
        else store_row_bf16<HC>(grow, half * HC, r, dh, 1.f);
      }
      // the next iteration's first wait uses the same (already completed) phase: do not flip here
      tcgen05_fence_before();
      __syncthreads();  // all rows read dQ before the next query block's MMAs (queued behind this commit) reuse it
      MDT_PROF(7)  // dQ read-out + barrier
    }
    phase ^= 1;  // the last commit of the item has been consumed by the wait above
    // dK / dV: rows = keys; threads 0-127 write dK, threads 128-255 write dV
#pragma unroll 1
    for (int kb = 0; kb < NB; ++kb) {
      const int key = kb * kQB + row;
      __nv_bfloat16* grow = dqkv + (static_cast<long long>(b) * T + key) * rs + ((1 + half) * H + h) * dh;
      const uint32_t tacc = tKV + kb * 2 * DP + half * DP;
      uint32_t r[DP];
      tmem_ld_cols<DP>(tacc + lane_addr, r);
      if constexpr (kBulkOut) stage_row_bf16<DP, DP>(sP + (1 + half) * kBlkBytes, row, 0, r, dh);
      else store_row_bf16<DP>(grow, 0, r, dh, 1.f);
    }
    if constexpr (kBulkOut) fence_proxy_async_smem();
    tcgen05_fence_before();
    __syncthreads();  // accumulators and tiles of this item are dead; the next item may overwrite them
    tcgen05_fence_after();
    if constexpr (kBulkOut) {
      if (warp == kLoadWarp) {
        const int lane = tid & 31;
        const int rb_per_box = (dh == DP) ? kQB / 8 : 1;
        const int per_tile = (kQB / 8) / rb_per_box, chunks = dh / 8;
        const int rb0 = static_cast<int>((static_cast<long long>(b) * T) / 8);
        for (int i = lane; i < 3 * per_tile; i += 32) {
          const int tile = i / per_tile, rblk = (i - tile * per_tile) * rb_per_box;
          tma_store_4d(&tm_dqkv, sP + tile * kBlkBytes + rblk * TT::ROWBLK, 0, 0, (tile * H + h) * chunks, rb0 + rblk);
        }
        bulk_commit_group();
      }
    }
    MDT_PROF(8)  // dK / dV read-out + barrier
  }
#ifdef MDT_ATTN_PROF
  if (blockIdx.x == 0 && (tid == 0 || tid == 200)) {
    float* dst = const_cast<float*>(lse) + static_cast<long long>(nitems) * T + (tid == 0 ? 0 : 16);
    for (int i = 0; i < 9; ++i) dst[i] = static_cast<float>(pt[i]);
  }
#endif
#undef MDT_PROF
  if constexpr (kBulkOut) {
    if (warp == kLoadWarp) bulk_wait_all();
  }
  if (warp == 0) tmem_dealloc<512>(tmem);
}

// ------------------------------------------------------------------------------------------------------------
// host dispatch (called from attention.cu)
// ------------------------------------------------------------------------------------------------------------
static bool attn_tma() {  // MDT_ATTN_TMA=0: cp.async tile fill (A/B switch)
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MDT_ATTN_TMA");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

template <int DP, int TK>
static int launch_fwd(const void* qkv, void* out, float* lse, int B, int T, int H, int dh, float scale,
                      cudaStream_t st) {
  const int smem = fwd_v_offset(DP, TK) + TK * DP * 2 + 4 * kQB * 4 + 64;
  static bool set = false;
  if (!set) {
    if (cudaFuncSetAttribute(attn_tc_fwd_kernel<DP, TK, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) !=
            cudaSuccess ||
        cudaFuncSetAttribute(attn_tc_fwd_kernel<DP, TK, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) !=
            cudaSuccess)
      return MDT_ERR_CUDA;
    set = true;
  }
  const dim3 grid(T / kQB, B * H);
  const int threads = kQB * fwd_tpr(TK);
  alignas(64) CUtensorMap tm;
  // measured (B200, us, TMA vs cp.async): T=128 d_h=72 95 vs 111; T=256 d_h=32 227 vs 199; T=256 d_h=72 186 vs 174 -
  // with 16-byte inner boxes TMA itself is the slower copy, it only wins where three CTAs per SM hide it
  const bool tma = attn_tma() && TK == 128 && (reinterpret_cast<uintptr_t>(qkv) & 15) == 0;
  if (tma) {
    const int rc = make_token_tile_tmap(&tm, qkv, static_cast<unsigned long long>(B) * T, 3ull * H * dh, dh / 8,
                                        dh == DP ? kQB / 8 : 1);
    if (rc != MDT_OK) return rc;
    attn_tc_fwd_kernel<DP, TK, true><<<grid, threads, smem, st>>>(static_cast<const __nv_bfloat16*>(qkv), tm,
                                                                  static_cast<__nv_bfloat16*>(out), lse, T, H, dh,
                                                                  scale);
  } else {
    memset(&tm, 0, sizeof(tm));
    attn_tc_fwd_kernel<DP, TK, false><<<grid, threads, smem, st>>>(static_cast<const __nv_bfloat16*>(qkv), tm,
                                                                   static_cast<__nv_bfloat16*>(out), lse, T, H, dh,
                                                                   scale);
  }
  return cudaGetLastError() == cudaSuccess ? MDT_OK : MDT_ERR_CUDA;
}
template <int DP, int NB>
static int launch_bwd(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, int B, int H,
                      int dh, float scale, cudaStream_t st) {
  constexpr int kSetTiles = (NB == 1) ? 4 : 5;
  constexpr int T = NB * kQB;
  const int smem = 2 * kSetTiles * NB * kQB * DP * 2 + 2 * kQB * kQB * 2 + 2 * kQB * 4 + 64;
  static bool set = false;
  static int sms = 0;
  if (!set) {
    if (cudaFuncSetAttribute(attn_tc_bwd_kernel<DP, NB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) !=
            cudaSuccess ||
        cudaFuncSetAttribute(attn_tc_bwd_kernel<DP, NB, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) !=
            cudaSuccess)
      return MDT_ERR_CUDA;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = kNumSMsDefault;
    set = true;
  }
  const int nitems = B * H;
  const int grid = nitems < sms ? nitems : sms;
  alignas(64) CUtensorMap tq, td, to, tg;
  const bool tma = attn_tma() && ((reinterpret_cast<uintptr_t>(qkv) | reinterpret_cast<uintptr_t>(out) |
                                   reinterpret_cast<uintptr_t>(dout) | reinterpret_cast<uintptr_t>(dqkv)) & 15) == 0;
  if (tma) {
    const unsigned long long rows = static_cast<unsigned long long>(B) * T;
    const unsigned rb = dh == DP ? kQB / 8 : 1;
    int rc = make_token_tile_tmap(&tq, qkv, rows, 3ull * H * dh, dh / 8, rb);
    if (rc == MDT_OK) rc = make_token_tile_tmap(&td, dout, rows, 1ull * H * dh, dh / 8, rb);
    if (rc == MDT_OK) rc = make_token_tile_tmap(&to, out, rows, 1ull * H * dh, dh / 8, rb);
    if (rc == MDT_OK) rc = make_token_tile_tmap(&tg, dqkv, rows, 3ull * H * dh, dh / 8, rb);
    if (rc != MDT_OK) return rc;
    attn_tc_bwd_kernel<DP, NB, true><<<grid, kBwdThreads, smem, st>>>(
        static_cast<const __nv_bfloat16*>(qkv), static_cast<const __nv_bfloat16*>(out),
        static_cast<const __nv_bfloat16*>(dout), tq, td, to, tg, lse, static_cast<__nv_bfloat16*>(dqkv), H, dh,
        scale, nitems);
  } else {
    memset(&tq, 0, sizeof(tq));
    attn_tc_bwd_kernel<DP, NB, false><<<grid, kBwdThreads, smem, st>>>(
        static_cast<const __nv_bfloat16*>(qkv), static_cast<const __nv_bfloat16*>(out),
        static_cast<const __nv_bfloat16*>(dout), tq, tq, tq, tq, lse, static_cast<__nv_bfloat16*>(dqkv), H, dh,
        scale, nitems);
  }
  return cudaGetLastError() == cudaSuccess ? MDT_OK : MDT_ERR_CUDA;
}

// returns MDT_ERR_UNSUPPORTED when the shape is outside the tcgen05 kernels' range (caller falls back to the
// mma.sync kernels, which handle any T / head_dim <= 80)
int attention_tc_fwd(const void* qkv, void* out, float* lse, int B, int T, int H, int dh, float scale, cudaStream_t st) {
  if (dh % 8) return MDT_ERR_UNSUPPORTED;
  const int dp = dh <= 32 ? 32 : (dh <= 64 ? 64 : (dh <= 80 ? 80 : 0));
  if (T == 128) {
    if (dp == 32) return launch_fwd<32, 128>(qkv, out, lse, B, T, H, dh, scale, st);
    if (dp == 64) return launch_fwd<64, 128>(qkv, out, lse, B, T, H, dh, scale, st);
    if (dp == 80) return launch_fwd<80, 128>(qkv, out, lse, B, T, H, dh, scale, st);
  } else if (T == 256) {
    if (dp == 32) return launch_fwd<32, 256>(qkv, out, lse, B, T, H, dh, scale, st);
    if (dp == 64) return launch_fwd<64, 256>(qkv, out, lse, B, T, H, dh, scale, st);
    if (dp == 80) return launch_fwd<80, 256>(qkv, out, lse, B, T, H, dh, scale, st);
  }
  return MDT_ERR_UNSUPPORTED;
}
int attention_tc_bwd(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, int B, int T,
                     int H, int dh, float scale, cudaStream_t st) {
  if (dh % 8) return MDT_ERR_UNSUPPORTED;
  const int dp = dh <= 32 ? 32 : (dh <= 64 ? 64 : (dh <= 80 ? 80 : 0));
  if (T == 128) {
    if (dp == 32) return launch_bwd<32, 1>(qkv, out, dout, lse, dqkv, B, H, dh, scale, st);
    if (dp == 64) return launch_bwd<64, 1>(qkv, out, dout, lse, dqkv, B, H, dh, scale, st);
    if (dp == 80) return launch_bwd<80, 1>(qkv, out, dout, lse, dqkv, B, H, dh, scale, st);
  } else if (T == 256) {
    if (dp == 32) return launch_bwd<32, 2>(qkv, out, dout, lse, dqkv, B, H, dh, scale, st);
  }
  return MDT_ERR_UNSUPPORTED;
}

}  // namespace mdt
