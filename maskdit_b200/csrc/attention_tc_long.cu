// tcgen05 attention for the longer MaskDiT sequences (T = NB * 128: the 512-px configs run the encoder at T = 512 and
// the decoder at T = 1024; also T = 256 with head_dim 72 in the backward).  Same tile layout / descriptor conventions
// as attention_tc.cu; what changes is that the T x T score matrix no longer fits TMEM, so:
//
//   forward : CTA = (128 queries of one (b,h)), all of K and V resident in smem, key blocks of 128 walked TWICE:
//             pass 1 computes the row maximum (S blocks double-buffered in TMEM), pass 2 recomputes each S block,
//             forms P = exp(S - max) and accumulates O += P V in TMEM without any rescaling of O.  The extra Q K^T
//             costs 50 % more (cheap) MMA work and removes the TMEM read-modify-write of an online softmax.
//   backward: two kernels without atomics (the accumulators of both operands of a (query block, key block) pair do not
//             fit TMEM together): dQ kernel, CTA = query block, streams K/V blocks; dK/dV kernel, CTA = key block,
//             streams Q/dO blocks.  Streamed blocks are double-buffered with cp.async.  delta = rowsum(dO * O) comes
//             from a small pre-kernel.
#include <stdlib.h>

#include "attention_tc.cuh"
#include "../../include/maskdit_b200.h"

namespace mdt {

constexpr int kLT = 256;  // threads per CTA: two per query row (row = tid & 127, column half = tid >> 7)
constexpr float kLog2e = 1.4426950408889634f;

MDT_DEVINL void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
MDT_DEVINL void cp_async_wait_group() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// ------------------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------------------
template <int DP, int NB>
__global__ void __launch_bounds__(kLT, 1)
attn_tcl_fwd_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, float* __restrict__ lse,
                    int H, int dh, float scale) {
  using TT = TokTile<DP>;
  constexpr int T = NB * kQB;
  constexpr int kPBlk = (kQB / 8) * 128;
  constexpr uint32_t RB = TT::ROWBLK, kBlkBytes = 16 * RB;
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t sQ = smem_u32(smem), sK = sQ + kQB * DP * 2, sV = sK + T * DP * 2, sP = sV + T * DP * 2;
  float* s_red = reinterpret_cast<float*>(smem + (kQB + 2 * T) * DP * 2 + kQB * kQB * 2);  // [2][128]
  uint64_t* bar = reinterpret_cast<uint64_t*>(s_red + 2 * kQB);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);

  const int tid = threadIdx.x, warp = tid >> 5, row = tid & (kQB - 1), half = tid >> 7;
  const int b = blockIdx.y / H, h = blockIdx.y % H, q0 = blockIdx.x * kQB;
  const long long rs = 3LL * H * dh;
  const __nv_bfloat16* base = qkv + static_cast<long long>(b) * T * rs;
  if (warp == 0) tmem_alloc<512>(tmem_slot);
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  TT::load(sQ, base + q0 * rs + h * dh, rs, kQB, dh);
  TT::load(sK, base + (H + h) * dh, rs, T, dh);
  TT::load(sV, base + (2 * H + h) * dh, rs, T, dh);
  cp_async_wait_all();
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tS = tmem, tO = tmem + 256;  // S blocks ping-pong between columns 0..127 and 128..255
  const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) * 32) << 16;
  const float sl = scale * kLog2e;
  uint32_t phase = 0;

  // ---- pass 1: row maximum over all key blocks ----
  if (tid == 0) {
    mma_kk<DP>(tS, sQ, sK, kQB, false);
    umma_commit(bar);
  }
  float m = -INFINITY;
  for (int j = 0; j < NB; ++j) {
    mbar_wait(bar, phase);
    phase ^= 1;
    tcgen05_fence_after();
    if (tid == 0 && j + 1 < NB) {  // next block into the other S buffer (its readers finished before the last sync)
      mma_kk<DP>(tS + ((j + 1) & 1) * 128, sQ, sK + (j + 1) * kBlkBytes, kQB, false);
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
    mma_kk<DP>(tS, sQ, sK, kQB, false);
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
      const uint32_t idesc = make_idesc_bf16(kQB, DP, 0, 1);  // A = P (K-major), B = V block (token-contracted)
#pragma unroll
      for (int k = 0; k < kQB / 16; ++k)
        umma_bf16(tO, make_smem_desc_nosw(sP + k * 256, 128, kPBlk),
                  make_smem_desc_nosw(sV + j * kBlkBytes + k * 2 * RB, RB, 128), idesc, (j > 0 || k > 0) ? 1u : 0u);
      if (j + 1 < NB) mma_kk<DP>(tS + ((j + 1) & 1) * 128, sQ, sK + (j + 1) * kBlkBytes, kQB, false);
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

// ------------------------------------------------------------------------------------------------------------
// delta[b,h,t] = sum_d dO[b,t,h,d] * O[b,t,h,d]
// ------------------------------------------------------------------------------------------------------------
__global__ void attn_delta_kernel(const __nv_bfloat16* __restrict__ out, const __nv_bfloat16* __restrict__ dout,
                                  float* __restrict__ delta, int B, int T, int H, int dh) {
  const long long n = static_cast<long long>(B) * T * H;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int h = static_cast<int>(i % H);
    const long long bt = i / H;
    const __nv_bfloat16* o = out + bt * (H * dh) + h * dh;
    const __nv_bfloat16* d = dout + bt * (H * dh) + h * dh;
    float acc = 0.f;
    for (int c = 0; c < dh; c += 8) {
      const uint4 a = ldg128u_nc(o + c), g = ldg128u_nc(d + c);
      acc += bf16_lo(a.x) * bf16_lo(g.x) + bf16_hi(a.x) * bf16_hi(g.x) + bf16_lo(a.y) * bf16_lo(g.y) +
             bf16_hi(a.y) * bf16_hi(g.y) + bf16_lo(a.z) * bf16_lo(g.z) + bf16_hi(a.z) * bf16_hi(g.z) +
             bf16_lo(a.w) * bf16_lo(g.w) + bf16_hi(a.w) * bf16_hi(g.w);
    }
    const long long bb = bt / T, t = bt % T;
    delta[(bb * H + h) * T + t] = acc;
  }
}

// delta pre-kernel launch (also used by attention_sw_long.cu)
void launch_attn_delta(const void* out, const void* dout, float* delta, int B, int T, int H, int dh, cudaStream_t st) {
  const long long n = static_cast<long long>(B) * T * H;
  attn_delta_kernel<<<static_cast<int>((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16), 256, 0, st>>>(
      static_cast<const __nv_bfloat16*>(out), static_cast<const __nv_bfloat16*>(dout), delta, B, T, H, dh);
}

// P / dS of this thread's half row from the S / dP accumulators; writes sdS (and sP if WRITE_P)
template <bool WRITE_P>
MDT_DEVINL void softmax_bwd_half(uint32_t tS, uint32_t tdP, uint32_t lane_addr, int half, int row, uint32_t sP,
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
// backward, dQ: CTA = (query block, (b,h)); K / V blocks streamed (double-buffered)
// ------------------------------------------------------------------------------------------------------------
template <int DP, int NB>
__global__ void __launch_bounds__(kLT, 1)
attn_tcl_dq_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ dout,
                   const float* __restrict__ lse, const float* __restrict__ delta_g,
                   __nv_bfloat16* __restrict__ dqkv, int H, int dh, float scale) {
  using TT = TokTile<DP>;
  constexpr int T = NB * kQB;
  constexpr int kPBlk = (kQB / 8) * 128;
  constexpr uint32_t RB = TT::ROWBLK, kBlkBytes = 16 * RB;
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t sQ = smem_u32(smem), sdO = sQ + kBlkBytes, sKV = sdO + kBlkBytes;  // sKV: 2 x (K block | V block)
  const uint32_t sdS = sKV + 4 * kBlkBytes;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 6 * kBlkBytes + kQB * kQB * 2);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);

  const int tid = threadIdx.x, warp = tid >> 5, row = tid & (kQB - 1), half = tid >> 7;
  const int b = blockIdx.y / H, h = blockIdx.y % H, qb = blockIdx.x;
  const long long rs = 3LL * H * dh;
  const int HD = H * dh;
  const __nv_bfloat16* base = qkv + static_cast<long long>(b) * T * rs;
  if (warp == 0) tmem_alloc<512>(tmem_slot);
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  auto load_kv = [&](int kb) {
    const uint32_t dst = sKV + (kb & 1) * 2 * kBlkBytes;
    TT::load(dst, base + static_cast<long long>(kb) * kQB * rs + (H + h) * dh, rs, kQB, dh);
    TT::load(dst + kBlkBytes, base + static_cast<long long>(kb) * kQB * rs + (2 * H + h) * dh, rs, kQB, dh);
    cp_async_commit();
  };
  TT::load(sQ, base + static_cast<long long>(qb) * kQB * rs + h * dh, rs, kQB, dh);
  TT::load(sdO, dout + (static_cast<long long>(b) * T + qb * kQB) * HD + h * dh, HD, kQB, dh);
  load_kv(0);
  if (NB > 1) load_kv(1);
  const int q = qb * kQB + row;
  const float lsl = lse[(static_cast<long long>(b) * H + h) * T + q] * kLog2e;
  const float delta = delta_g[(static_cast<long long>(b) * H + h) * T + q];
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tS = tmem, tdP = tmem + 128, tdQ = tmem + 256;
  const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) * 32) << 16;
  const float sl = scale * kLog2e;
  uint32_t phase = 0;

  for (int kb = 0; kb < NB; ++kb) {
    if (kb + 1 < NB) cp_async_wait_group<1>(); else cp_async_wait_group<0>();
    fence_proxy_async_smem();
    tcgen05_fence_before();
    __syncthreads();
    const uint32_t sK = sKV + (kb & 1) * 2 * kBlkBytes, sV = sK + kBlkBytes;
    if (tid == 0) {
      tcgen05_fence_after();
      mma_kk<DP>(tS, sQ, sK, kQB, false);
      mma_kk<DP>(tdP, sdO, sV, kQB, false);
      umma_commit(bar);
    }
    mbar_wait(bar, phase);
    phase ^= 1;
    tcgen05_fence_after();
    softmax_bwd_half<false>(tS, tdP, lane_addr, half, row, 0, sdS, sl, lsl, delta, scale);
    fence_proxy_async_smem();
    tcgen05_fence_before();
    __syncthreads();
    if (tid == 0) {
      tcgen05_fence_after();
      const uint32_t idesc = make_idesc_bf16(kQB, DP, 0, 1);  // A = dS (K-major: keys contiguous), B = K block
#pragma unroll
      for (int k = 0; k < kQB / 16; ++k)
        umma_bf16(tdQ, make_smem_desc_nosw(sdS + k * 256, 128, kPBlk), make_smem_desc_nosw(sK + k * 2 * RB, RB, 128),
                  idesc, (kb > 0 || k > 0) ? 1u : 0u);
      umma_commit(bar);
    }
    mbar_wait(bar, phase);  // dS tile and this K/V buffer are free again
    phase ^= 1;
    tcgen05_fence_after();
    if (kb + 2 < NB) load_kv(kb + 2);
  }
  __nv_bfloat16* grow = dqkv + (static_cast<long long>(b) * T + q) * rs + h * dh;
  {
    constexpr int HC = DP / 2;
    uint32_t r[HC];
    tmem_ld_cols<HC>(tdQ + lane_addr + half * HC, r);
    store_row_bf16<HC>(grow, half * HC, r, dh, 1.f);
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) {
    tcgen05_fence_after();
    tmem_dealloc<512>(tmem);
  }
}

// ------------------------------------------------------------------------------------------------------------
// backward, dK / dV: CTA = (key block, (b,h)); Q / dO blocks streamed (double-buffered)
// ------------------------------------------------------------------------------------------------------------
template <int DP, int NB>
__global__ void __launch_bounds__(kLT, 1)
attn_tcl_dkv_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ dout,
                    const float* __restrict__ lse, const float* __restrict__ delta_g,
                    __nv_bfloat16* __restrict__ dqkv, int H, int dh, float scale) {
  using TT = TokTile<DP>;
  constexpr int T = NB * kQB;
  constexpr int kPBlk = (kQB / 8) * 128;
  constexpr uint32_t RB = TT::ROWBLK, kBlkBytes = 16 * RB;
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t sK = smem_u32(smem), sV = sK + kBlkBytes, sQO = sV + kBlkBytes;  // sQO: 2 x (Q block | dO block)
  const uint32_t sP = sQO + 4 * kBlkBytes, sdS = sP + kQB * kQB * 2;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 6 * kBlkBytes + 2 * kQB * kQB * 2);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);

  const int tid = threadIdx.x, warp = tid >> 5, row = tid & (kQB - 1), half = tid >> 7;
  const int b = blockIdx.y / H, h = blockIdx.y % H, kb = blockIdx.x;
  const long long rs = 3LL * H * dh;
  const int HD = H * dh;
  const __nv_bfloat16* base = qkv + static_cast<long long>(b) * T * rs;
  if (warp == 0) tmem_alloc<512>(tmem_slot);
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  auto load_qo = [&](int qb) {
    const uint32_t dst = sQO + (qb & 1) * 2 * kBlkBytes;
    TT::load(dst, base + static_cast<long long>(qb) * kQB * rs + h * dh, rs, kQB, dh);
    TT::load(dst + kBlkBytes, dout + (static_cast<long long>(b) * T + qb * kQB) * HD + h * dh, HD, kQB, dh);
    cp_async_commit();
  };
  TT::load(sK, base + static_cast<long long>(kb) * kQB * rs + (H + h) * dh, rs, kQB, dh);
  TT::load(sV, base + static_cast<long long>(kb) * kQB * rs + (2 * H + h) * dh, rs, kQB, dh);
  load_qo(0);
  if (NB > 1) load_qo(1);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tS = tmem, tdP = tmem + 128, tdK = tmem + 256, tdV = tmem + 256 + DP;
  const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) * 32) << 16;
  const float sl = scale * kLog2e;
  uint32_t phase = 0;

  for (int qb = 0; qb < NB; ++qb) {
    const int q = qb * kQB + row;
    const float lsl = lse[(static_cast<long long>(b) * H + h) * T + q] * kLog2e;
    const float delta = delta_g[(static_cast<long long>(b) * H + h) * T + q];
    if (qb + 1 < NB) cp_async_wait_group<1>(); else cp_async_wait_group<0>();
    fence_proxy_async_smem();
    tcgen05_fence_before();
    __syncthreads();
    const uint32_t sQ = sQO + (qb & 1) * 2 * kBlkBytes, sdO = sQ + kBlkBytes;
    if (tid == 0) {
      tcgen05_fence_after();
      mma_kk<DP>(tS, sQ, sK, kQB, false);    // S[query, key]
      mma_kk<DP>(tdP, sdO, sV, kQB, false);  // dP[query, key]
      umma_commit(bar);
    }
    mbar_wait(bar, phase);
    phase ^= 1;
    tcgen05_fence_after();
    softmax_bwd_half<true>(tS, tdP, lane_addr, half, row, sP, sdS, sl, lsl, delta, scale);
    fence_proxy_async_smem();
    tcgen05_fence_before();
    __syncthreads();
    if (tid == 0) {
      tcgen05_fence_after();
      const uint32_t idesc = make_idesc_bf16(kQB, DP, 1, 1);  // both operands token(query)-contracted
#pragma unroll
      for (int k = 0; k < kQB / 16; ++k) {
        umma_bf16(tdV, make_smem_desc_nosw(sP + k * 2 * kPBlk, kPBlk, 128),
                  make_smem_desc_nosw(sdO + k * 2 * RB, RB, 128), idesc, (qb > 0 || k > 0) ? 1u : 0u);
        umma_bf16(tdK, make_smem_desc_nosw(sdS + k * 2 * kPBlk, kPBlk, 128),
                  make_smem_desc_nosw(sQ + k * 2 * RB, RB, 128), idesc, (qb > 0 || k > 0) ? 1u : 0u);
      }
      umma_commit(bar);
    }
    mbar_wait(bar, phase);  // P / dS tiles and this Q/dO buffer are free again
    phase ^= 1;
    tcgen05_fence_after();
    if (qb + 2 < NB) load_qo(qb + 2);
  }
  {
    const int key = kb * kQB + row;
    __nv_bfloat16* grow = dqkv + (static_cast<long long>(b) * T + key) * rs + ((1 + half) * H + h) * dh;
    uint32_t r[DP];
    tmem_ld_cols<DP>((half ? tdV : tdK) + lane_addr, r);
    store_row_bf16<DP>(grow, 0, r, dh, 1.f);
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) {
    tcgen05_fence_after();
    tmem_dealloc<512>(tmem);
  }
}

// ------------------------------------------------------------------------------------------------------------
// host dispatch
// ------------------------------------------------------------------------------------------------------------
template <typename K>
static int set_smem(K kern, int smem) {
  return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) == cudaSuccess ? MDT_OK
                                                                                                      : MDT_ERR_CUDA;
}

template <int DP, int NB>
static int launch_long_fwd(const void* qkv, void* out, float* lse, int B, int H, int dh, float scale, cudaStream_t st) {
  constexpr int T = NB * kQB;
  constexpr int smem = (kQB + 2 * T) * DP * 2 + kQB * kQB * 2 + 2 * kQB * 4 + 64;
  if constexpr (smem > 232448) {
    return MDT_ERR_UNSUPPORTED;
  } else {
    auto kern = attn_tcl_fwd_kernel<DP, NB>;
    static bool set = false;
    if (!set) {
      if (int rc = set_smem(kern, smem)) return rc;
      set = true;
    }
    kern<<<dim3(NB, B * H), kLT, smem, st>>>(static_cast<const __nv_bfloat16*>(qkv), static_cast<__nv_bfloat16*>(out),
                                             lse, H, dh, scale);
    return cudaGetLastError() == cudaSuccess ? MDT_OK : MDT_ERR_CUDA;
  }
}

template <int DP, int NB>
static int launch_long_bwd(const void* qkv, const void* out, const void* dout, const float* lse, float* delta,
                           void* dqkv, int B, int H, int dh, float scale, cudaStream_t st) {
  constexpr int T = NB * kQB;
  constexpr int kBlk = kQB * DP * 2;
  constexpr int smem_dq = 6 * kBlk + kQB * kQB * 2 + 64;
  constexpr int smem_dkv = 6 * kBlk + 2 * kQB * kQB * 2 + 64;
  auto k_dq = attn_tcl_dq_kernel<DP, NB>;
  auto k_dkv = attn_tcl_dkv_kernel<DP, NB>;
  static bool set = false;
  if (!set) {
    if (int rc = set_smem(k_dq, smem_dq)) return rc;
    if (int rc = set_smem(k_dkv, smem_dkv)) return rc;
    set = true;
  }
  launch_attn_delta(out, dout, delta, B, T, H, dh, st);
  k_dq<<<dim3(NB, B * H), kLT, smem_dq, st>>>(static_cast<const __nv_bfloat16*>(qkv),
                                              static_cast<const __nv_bfloat16*>(dout), lse, delta,
                                              static_cast<__nv_bfloat16*>(dqkv), H, dh, scale);
  k_dkv<<<dim3(NB, B * H), kLT, smem_dkv, st>>>(static_cast<const __nv_bfloat16*>(qkv),
                                                static_cast<const __nv_bfloat16*>(dout), lse, delta,
                                                static_cast<__nv_bfloat16*>(dqkv), H, dh, scale);
  return cudaGetLastError() == cudaSuccess ? MDT_OK : MDT_ERR_CUDA;
}

int attention_tc_long_fwd(const void* qkv, void* out, float* lse, int B, int T, int H, int dh, float scale,
                          cudaStream_t st) {
  if (dh % 8) return MDT_ERR_UNSUPPORTED;
  const int dp = dh <= 32 ? 32 : (dh <= 64 ? 64 : (dh <= 80 ? 80 : 0));
  if (T == 512) {
    if (dp == 32) return launch_long_fwd<32, 4>(qkv, out, lse, B, H, dh, scale, st);
    if (dp == 64) return launch_long_fwd<64, 4>(qkv, out, lse, B, H, dh, scale, st);
    if (dp == 80) return launch_long_fwd<80, 4>(qkv, out, lse, B, H, dh, scale, st);
  } else if (T == 1024) {
    if (dp == 32) return launch_long_fwd<32, 8>(qkv, out, lse, B, H, dh, scale, st);
  }
  return MDT_ERR_UNSUPPORTED;
}

// `delta` = scratch [B,H,T] floats
int attention_tc_long_bwd(const void* qkv, const void* out, const void* dout, const float* lse, float* delta,
                          void* dqkv, int B, int T, int H, int dh, float scale, cudaStream_t st) {
  if (dh % 8) return MDT_ERR_UNSUPPORTED;
  const int dp = dh <= 32 ? 32 : (dh <= 64 ? 64 : (dh <= 80 ? 80 : 0));
#define MDT_LB(DPV, NBV) return launch_long_bwd<DPV, NBV>(qkv, out, dout, lse, delta, dqkv, B, H, dh, scale, st)
  if (T == 256) {
    if (dp == 64) MDT_LB(64, 2);
    if (dp == 80) MDT_LB(80, 2);
  } else if (T == 512) {
    if (dp == 32) MDT_LB(32, 4);
    if (dp == 64) MDT_LB(64, 4);
    if (dp == 80) MDT_LB(80, 4);
  } else if (T == 1024) {
    if (dp == 32) MDT_LB(32, 8);
    if (dp == 64) MDT_LB(64, 8);
    if (dp == 80) MDT_LB(80, 8);
  }
#undef MDT_LB
  return MDT_ERR_UNSUPPORTED;
}

}  // namespace mdt
