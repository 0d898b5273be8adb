// bf16 GEMM on the 5th-gen tensor cores (tcgen05.mma, accumulators in TMEM, operands staged by TMA).
//
//   out[M,N] (+)= sum_k A[m,k] * B[n,k]      fp32 accumulate
//
// One persistent CTA per SM, warp-specialised:
//   warp 0      TMA producer   (one thread)  global -> 128B-swizzled smem ring, mbarrier complete_tx
//   warp 1      MMA issuer     (one thread)  tcgen05.mma 128 x BLOCK_N x 16, tcgen05.commit -> barriers
//   warp 2      TMEM allocator
//   warps 4-11  epilogue       tcgen05.ld (lane group = warp % 4, column half = (warp-4)/4) -> fused epilogue
// Two TMEM accumulator stages so the epilogue of unit i overlaps the MMAs of unit i+1.
//
// Operand majors: "K-major" = contraction index contiguous in memory (activations [M,K], nn.Linear weights
// [N,K]); "MN-major" = the M/N index contiguous (used by dgrad: B = W[N,K] read as [K_out, N_contract]; and by
// wgrad: both operands are token-major [tokens, features] with the contraction over tokens).
//
// Scheduling: persistent tile loop (each CTA group takes whole output tiles, strided) for forward/dgrad; for the
// accumulate epilogue (wgrad, long-K small-output GEMMs) the K range is additionally cut into slices so that
// tiles x slices fills the machine, units are walked slice-major and partial tiles are reduced with fp32
// red.global.add.v4 in the (pipelined) epilogue.
//
// Reference ops this replaces: every nn.Linear of models/maskdit.py (timm Attention.qkv/proj, Mlp.fc1/fc2,
// adaLN_modulation, DecoderLayer.linear, TimestepEmbedder.mlp, LabelEmbedder) and their autograd backward.
#include "common.cuh"
#include "gemm.h"
#include "unit_sched.h"

#include <stdlib.h>

#include <mutex>

namespace mdt {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // 64 bf16 = 128 B = one swizzle-128B row
constexpr int UMMA_K = 16;
#ifndef MDT_EPI_WARPS
#define MDT_EPI_WARPS 8
#endif
// Epilogue warps per CTA (multiple of 4: warp w may only touch TMEM lanes 32*(w%4)..+31).  The 32-column chunks of a
// tile are dealt to the kNumEpiWarps/4 warps of a lane group round robin.  With more than 8 the register file is
// re-partitioned at kernel start (setmaxnreg: the producer/MMA warpgroup keeps 40 registers per thread).
constexpr int kNumEpiWarps = MDT_EPI_WARPS;
// L2 policy experiments (profiles/r02_experiments.md section 12): MDT_EPI_CS = 1 makes the epilogue's once-touched
// global traffic (outputs, residual / GELU operand reads) evict-first; MDT_TMA_HINT = 1 loads the operand tiles with
// the evict-last policy.
#ifndef MDT_EPI_CS
#define MDT_EPI_CS 0
#endif
#ifndef MDT_TMA_HINT
#define MDT_TMA_HINT 0
#endif
#if MDT_EPI_CS
#define MDT_STG128 stg128_cs
#define MDT_STG64 stg64_cs
#define MDT_LDG128 ldg128_cs
#else
#define MDT_STG128 stg128
#define MDT_STG64 stg64
#define MDT_LDG128 ldg128
#endif
static_assert(kNumEpiWarps % 4 == 0 && kNumEpiWarps >= 4 && kNumEpiWarps <= 16, "epilogue warps: 4, 8, 12 or 16");
constexpr int kNumThreads = 128 + kNumEpiWarps * 32;  // 384 (8 warps) / 512 (12) / 640 (16)
constexpr int kEpiRegs = kNumEpiWarps == 12 ? 152 : (kNumEpiWarps == 16 ? 104 : 0);  // setmaxnreg.inc target

template <int BLOCK_N, int CG>
struct GemmCfg {
  // CG = 1: one CTA per 128 x BLOCK_N tile.  CG = 2: an SM pair per 256 x BLOCK_N tile (tcgen05 cta_group::2):
  // each CTA stages its own 128 rows of A and HALF of the B tile, which halves the smem fill+read traffic per
  // MMA flop (a single-CTA 128x256 tile needs 187 B/clk of smem bandwidth at full tensor rate, the SM has 128).
  static constexpr int kBRows = BLOCK_N / CG;
  static constexpr int kABytes = BLOCK_M * BLOCK_K * 2;
  static constexpr int kBBytes = kBRows * BLOCK_K * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStagingBytes = kNumEpiWarps * 32 * 36 * 4;  // per-warp 32x36 fp32 transpose tiles
  static constexpr int kFixed = kStagingBytes + 1024 /*align slack*/ + 256 /*barriers*/;
  static constexpr int kStagesFit = (232448 - kFixed) / kStageBytes;
  static constexpr int kStages = kStagesFit > 8 ? 8 : kStagesFit;
  static constexpr int kTmemCols = (2 * BLOCK_N > 256) ? 512 : 256;
  static constexpr int kSmemBytes = kStages * kStageBytes + kFixed;
  static_assert(kStages >= 3, "pipeline too shallow");
};

// ---- fused epilogue ---------------------------------------------------------------------------------------
// tcgen05.ld hands each thread one accumulator ROW (32 consecutive fp32 columns).  Writing rows straight to global
// makes every warp store touch 32 different 128-byte lines (ncu v1: 32 sectors/request, tensor pipe 18-50 % busy).
// The 32x32 chunk is therefore transposed through a per-warp smem staging tile (row stride 36 words: STS.128 by
// row owners and LDS.128 by (row, 4-column) owners are both conflict free per quarter warp).  In the second phase
// lane -> (row = lane/8 + 4i, 4 columns = lane%8): one warp instruction covers 4 rows x 128 contiguous bytes (fp32)
// or 64 bytes (bf16), and bias / gate / residual / aux are read with the same coalesced vector pattern.
// The per-epilogue loops are separate template instances selected by ONE switch per chunk, so a launch only ever
// touches the instructions of its own epilogue (ncu v3 showed the epilogue instruction-fetch bound).
#ifdef MDT_GEMM_PROF
__device__ float g_gemm_prof[12];
#endif
constexpr int kStgStride = 36;
constexpr int kStgFloats = 32 * kStgStride;

MDT_DEVINL uint2 pack4_bf16(float4 v) { return make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w)); }
MDT_DEVINL float4 unpack4_bf16(uint2 u) { return make_float4(bf16_lo(u.x), bf16_hi(u.x), bf16_lo(u.y), bf16_hi(u.y)); }

// Global operands of one chunk (this lane's 4 columns x 8 row groups) are fetched for all 8 row groups BEFORE any
// dependent store is issued: issue is in order, so a load -> store dependency inside the row loop would expose the
// full global latency 8 times per chunk (ncu r01: the K=1152 proj GEMM ran at 18 % tensor-pipe activity because of
// exactly that).  Measured and rejected on top of this (A/B on one box, in-step timing, tools/gemm_shapes_step.py):
//   - a register-level software pipeline (row group i refilled with the next chunk's values right after its store):
//     2x SLOWER on the K=1152 gate+residual GEMM - the in-flight loads share scoreboards with the following LDS;
//   - cp.async.bulk.prefetch.L2 of the next tile's operand rows: 3-4 % slower on the dGELU / gate+residual GEMMs;
//   - issuing the chunk's loads above the tcgen05.ld (as done now): neutral.
// The operand latency is therefore not what separates these epilogues (1000-1200 TF/s in-step) from the operand-free
// ones (1250-1390): it is their instruction count on two warps per scheduler.
struct EpiCoord {
  int row_base, nrows, col;  // col = this lane's first column
  bool valid;                // nrows > 0 && col + 4 <= N
};
MDT_DEVINL EpiCoord make_coord(const GemmParams& p, int row_base, int nrows, int col) {
  return EpiCoord{row_base, nrows, col, nrows > 0 && col + 4 <= p.N};
}

template <int EPI>
struct EpiOps {
  float4 bias4, gate4;
  bool gate_uniform;
  float4 res[(EPI == EPI_GATE_RESID || EPI == EPI_STORE) ? 8 : 1];
  uint2 auxv[EPI == EPI_DGELU ? 8 : 1];
};

template <int EPI>
MDT_DEVINL void epi_load(const GemmParams& p, const EpiCoord& c, int lane, EpiOps<EPI>& o) {
  if (!c.valid) return;
  const int rsub = lane >> 3;
  const size_t row0 = static_cast<size_t>(c.row_base + rsub);
  o.bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (EPI != EPI_ATOMIC && p.bias) o.bias4 = ldg128_nc(gaddr(p.bias + c.col));
  if constexpr (EPI == EPI_GATE_RESID) {
    const int b0 = c.row_base / p.rows_per_group, b1 = (c.row_base + c.nrows - 1) / p.rows_per_group;
    o.gate_uniform = b0 == b1;
    o.gate4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (o.gate_uniform) o.gate4 = ldg128_nc(gaddr(p.gate + static_cast<size_t>(b0) * p.ld_gate + c.col));
  }
  if constexpr (EPI == EPI_GATE_RESID || EPI == EPI_STORE) {
    if (EPI == EPI_GATE_RESID || p.resid) {
      const uint64_t a_res = gaddr(p.resid) + (row0 * p.ld_resid + c.col) * 4, s_res = 16ull * p.ld_resid;
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (4 * i + rsub < c.nrows) o.res[i] = MDT_LDG128(a_res + i * s_res);
    }
  }
  if constexpr (EPI == EPI_DGELU) {
    const uint64_t a_aux = gaddr(p.aux) + (row0 * p.ld_aux + c.col) * 2, s_aux = 8ull * p.ld_aux;
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (4 * i + rsub < c.nrows) o.auxv[i] = ldg64_nc(a_aux + i * s_aux);
  }
}

// chunk `c` (valid for this lane): transpose-read, fused arithmetic, coalesced stores
template <int EPI>
MDT_DEVINL void epilogue_chunk(const GemmParams& p, uint32_t stg, const EpiCoord& c, int lane, const EpiOps<EPI>& o,
                               float4& cs_out) {
  const int rsub = lane >> 3, c4 = (lane & 7) * 4;
  const float4 bias4 = o.bias4;
  const size_t row0 = static_cast<size_t>(c.row_base + rsub);
  const int osz = (EPI == EPI_ATOMIC || EPI == EPI_GATE_RESID || (EPI == EPI_STORE && p.out_fp32)) ? 4 : 2;
  const uint64_t a_out = gaddr(p.out) + (row0 * p.ldo + c.col) * osz;
  const uint64_t a_aux = gaddr(p.aux) + (row0 * p.ld_aux + c.col) * 2;
  const uint64_t s_out = 4ull * p.ldo * osz, s_aux = 8ull * p.ld_aux;
  const uint32_t sp = stg + (rsub * kStgStride + c4) * 4;
  float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);  // EPI_DGELU: column sums of this lane's outputs (bias gradient)
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if (4 * i + rsub >= c.nrows) break;
    float4 v = lds128(sp + i * (4 * kStgStride * 4));
    const uint64_t ao = a_out + i * s_out;
    if constexpr (EPI == EPI_ATOMIC) {
      red_add_v4(ao, v);
    } else {
      v.x += bias4.x, v.y += bias4.y, v.z += bias4.z, v.w += bias4.w;
      if constexpr (EPI == EPI_STORE) {
        if (p.resid) v.x += o.res[i].x, v.y += o.res[i].y, v.z += o.res[i].z, v.w += o.res[i].w;
        if (p.act == ACT_SILU) v = make_float4(silu(v.x), silu(v.y), silu(v.z), silu(v.w));
        if (p.out_fp32) MDT_STG128(ao, v); else MDT_STG64(ao, pack4_bf16(v));
      } else if constexpr (EPI == EPI_GELU) {
        // pre-activation is rounded to bf16 first (as a bf16 nn.Linear output would be), GELU on the rounded value
        // (storing gelu'(h) here instead, so that the backward epilogue only multiplies, was measured: the forward GEMM
        // went 286 -> 333 us, the dGELU GEMM 333 -> 300 us - a net loss, profiles/r02_experiments.md section 7)
        const uint2 pre = pack4_bf16(v);
        if (p.aux) MDT_STG64(a_aux + i * s_aux, pre);
        const float4 h = unpack4_bf16(pre);
        MDT_STG64(ao, pack4_bf16(make_float4(gelu_tanh(h.x), gelu_tanh(h.y), gelu_tanh(h.z), gelu_tanh(h.w))));
      } else if constexpr (EPI == EPI_GATE_RESID) {
        if (p.aux) MDT_STG64(a_aux + i * s_aux, pack4_bf16(v));
        float4 g = o.gate4;
        if (!o.gate_uniform) {
          const size_t b = (row0 + 4 * i) / p.rows_per_group;
          g = ldg128_nc(gaddr(p.gate + b * p.ld_gate + c.col));
        }
        const float4 r = o.res[i];  // may alias `out` (in-place residual update): each element is read before written
        MDT_STG128(ao, make_float4(fmaf(g.x, v.x, r.x), fmaf(g.y, v.y, r.y), fmaf(g.z, v.z, r.z), fmaf(g.w, v.w, r.w)));
      } else if constexpr (EPI == EPI_DGELU) {
        const float4 h = unpack4_bf16(o.auxv[i]);
        const uint2 ov = pack4_bf16(make_float4(v.x * gelu_tanh_grad(h.x), v.y * gelu_tanh_grad(h.y),
                                                v.z * gelu_tanh_grad(h.z), v.w * gelu_tanh_grad(h.w)));
        MDT_STG64(ao, ov);
        const float4 r = unpack4_bf16(ov);  // sum what was stored (what a separate column-sum pass would read back)
        cs.x += r.x, cs.y += r.y, cs.z += r.z, cs.w += r.w;
      }
    }
  }
  if constexpr (EPI == EPI_DGELU) cs_out = cs;
}

// ragged right edge (N % 4 != 0 inside this 4-column group): element-wise, same arithmetic, cold path
__device__ __noinline__ void epilogue_ragged(const GemmParams& p, uint32_t stg, int row_base, int nrows, int col,
                                             int lane) {
  const int rsub = lane >> 3, c4 = (lane & 7) * 4;
  for (int i = 0; i < 8; ++i) {
    const int rr = 4 * i + rsub;
    if (rr >= nrows) break;
    const size_t row = static_cast<size_t>(row_base + rr);
    for (int j = 0; j < 4 && col + j < p.N; ++j) {
      float v = lds32(stg + (rr * kStgStride + c4 + j) * 4);
      const int c = col + j;
      if (p.epi == EPI_ATOMIC) {
        atomicAdd(reinterpret_cast<float*>(p.out) + row * p.ldo + c, v);
        continue;
      }
      if (p.bias) v += p.bias[c];
      __nv_bfloat16* o16 = reinterpret_cast<__nv_bfloat16*>(p.out) + row * p.ldo + c;
      float* o32 = reinterpret_cast<float*>(p.out) + row * p.ldo + c;
      __nv_bfloat16* aux = reinterpret_cast<__nv_bfloat16*>(p.aux) + row * p.ld_aux + c;
      switch (p.epi) {
        case EPI_STORE:
          if (p.resid) v += p.resid[row * p.ld_resid + c];
          if (p.act == ACT_SILU) v = silu(v);
          if (p.out_fp32) *o32 = v; else *o16 = __float2bfloat16_rn(v);
          break;
        case EPI_GELU: {
          const __nv_bfloat16 pre = __float2bfloat16_rn(v);
          if (p.aux) *aux = pre;
          *o16 = __float2bfloat16_rn(gelu_tanh(__bfloat162float(pre)));
        } break;
        case EPI_GATE_RESID:
          if (p.aux) *aux = __float2bfloat16_rn(v);
          *o32 = fmaf(p.gate[(row / p.rows_per_group) * p.ld_gate + c], v, p.resid[row * p.ld_resid + c]);
          break;
        case EPI_DGELU: {
          const __nv_bfloat16 r = __float2bfloat16_rn(v * gelu_tanh_grad(__bfloat162float(*aux)));
          *o16 = r;
          if (p.colsum) atomicAdd(p.colsum + c, __bfloat162float(r));
        } break;
        default: break;
      }
    }
  }
}

// The epilogue warps' whole persistent loop for ONE epilogue kind (one switch per launch: a launch only ever executes
// the instructions of its own epilogue; ncu v3 showed the epilogue instruction-fetch bound).  The chunk loop is
// deliberately rolled and the TMEM read single-buffered: unrolling / prefetching the next tcgen05.ld was measured
// SLOWER on B200 (12.3k vs 8.8k cycles per 128x256 bf16 tile, tools/probe_gemm2.py).
template <int EPI, int BLOCK_N, int CG>
MDT_DEVINL void epilogue_loop(const GemmParams& p, UnitSched& sched, uint64_t* tmem_full_bar,
                              uint64_t* tmem_empty_bar, uint32_t tmem_base, uint32_t stg, int warp, int cta_rank,
                              int lane) {
  constexpr int TILE_M = BLOCK_M * CG;
  constexpr int kParts = kNumEpiWarps / 4;  // warps sharing one TMEM lane group
  constexpr int kTileChunks = BLOCK_N / 32;
  constexpr bool kHasOps = EPI == EPI_STORE || EPI == EPI_GELU || EPI == EPI_GATE_RESID || EPI == EPI_DGELU;
  const int lane_group = warp & 3;       // tcgen05.ld: warp w may touch TMEM lanes 32*(w%4)..+31
  const int part = (warp - 4) >> 2;      // this warp takes chunks part, part + kParts, ... of the tile
  const int lcol = (lane & 7) * 4;
  auto tile_coord = [&](int mt, int nt, int& row_base, int& nrows, int& col_base) {
    row_base = mt * TILE_M + cta_rank * BLOCK_M + lane_group * 32;
    nrows = p.M - row_base;
    nrows = nrows > 32 ? 32 : nrows;
    col_base = nt * BLOCK_N;
  };
  int as = 0;
  uint32_t aphase = 0;
#ifdef MDT_GEMM_PROF  // phase cycle counters of one epilogue warp (tools/gemm_phase_prof.py)
  long long pt[5] = {0, 0, 0, 0, 0}, t_prev = clock64();
  int n_tiles = 0;
#define MDT_GPROF(i) { const long long t_now = clock64(); pt[i] += t_now - t_prev; t_prev = t_now; }
#else
#define MDT_GPROF(i)
#endif
  bool have = sched.next();
  while (have) {
    int row_base, nrows, col_base;
    tile_coord(sched.m_tile(), sched.n_tile(), row_base, nrows, col_base);
    have = sched.next();
    while (!mbar_try_wait(&tmem_full_bar[as], aphase)) {
    }
    tcgen05_fence_after();
    MDT_GPROF(0)  // waiting for the accumulator
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(lane_group * 32) << 16) + as * BLOCK_N;
#pragma unroll 1
    for (int ci = part; ci < kTileChunks; ci += kParts) {
      const int col0 = col_base + ci * 32;
      const EpiCoord c = make_coord(p, row_base, nrows, col0 + lcol);
      EpiOps<EPI> ops;
      if constexpr (kHasOps) epi_load<EPI>(p, c, lane, ops);
      uint32_t rc[32];
      tmem_ld_32x32b_x32(taddr + ci * 32, rc);
      tcgen05_wait_ld();
      MDT_GPROF(1)  // operand load issue + tcgen05.ld
      if (nrows > 0 && col0 < p.N) {  // warp-uniform
#pragma unroll
        for (int q = 0; q < 8; ++q)
          sts128(stg + (lane * kStgStride + 4 * q) * 4, __uint_as_float(rc[4 * q]), __uint_as_float(rc[4 * q + 1]),
                 __uint_as_float(rc[4 * q + 2]), __uint_as_float(rc[4 * q + 3]));
        __syncwarp();
        MDT_GPROF(2)  // staging stores
        if constexpr (kHasOps || EPI == EPI_ATOMIC) {
          float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);
          if (c.valid) epilogue_chunk<EPI>(p, stg, c, lane, ops, cs);
          else if (c.col < p.N) epilogue_ragged(p, stg, row_base, nrows, c.col, lane);
          if constexpr (EPI == EPI_DGELU) {
            if (p.colsum) {  // launch-uniform: bias gradient = column sums of the stored tile (fused colsum pass)
              __syncwarp();  // lanes l, l^8, l^16, l^24 own the same 4 columns (different row groups)
              cs.x += __shfl_xor_sync(0xffffffffu, cs.x, 8), cs.y += __shfl_xor_sync(0xffffffffu, cs.y, 8);
              cs.z += __shfl_xor_sync(0xffffffffu, cs.z, 8), cs.w += __shfl_xor_sync(0xffffffffu, cs.w, 8);
              cs.x += __shfl_xor_sync(0xffffffffu, cs.x, 16), cs.y += __shfl_xor_sync(0xffffffffu, cs.y, 16);
              cs.z += __shfl_xor_sync(0xffffffffu, cs.z, 16), cs.w += __shfl_xor_sync(0xffffffffu, cs.w, 16);
              if (lane < 8 && c.valid) red_add_v4(gaddr(p.colsum + c.col), cs);
            }
          }
        }
        __syncwarp();
        MDT_GPROF(3)  // transposed read, fused math, global stores
      }
    }
    tcgen05_fence_before();
    __syncwarp();
    if (lane == 0) {
      if constexpr (CG == 2) mbar_arrive_cluster(&tmem_empty_bar[as], 0);
      else mbar_arrive(&tmem_empty_bar[as]);
    }
    if (++as == 2) as = 0, aphase ^= 1;
    MDT_GPROF(4)  // release of the accumulator stage
#ifdef MDT_GEMM_PROF
    ++n_tiles;
#endif
  }
#ifdef MDT_GEMM_PROF
  if (blockIdx.x == 0 && warp == 5 && lane == 0) {
    for (int i = 0; i < 5; ++i) g_gemm_prof[i] = static_cast<float>(pt[i]);
    g_gemm_prof[5] = static_cast<float>(n_tiles);
  }
#endif
#undef MDT_GPROF
}

template <int BLOCK_N, bool A_MN, bool B_MN, int CG>
__global__ void __launch_bounds__(kNumThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const GemmParams p) {
  using Cfg = GemmCfg<BLOCK_N, CG>;
  constexpr int kStages = Cfg::kStages;
  constexpr int TILE_M = BLOCK_M * CG;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_tiles = smem;
  float* staging = reinterpret_cast<float*>(smem + kStages * Cfg::kStageBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * Cfg::kStageBytes + Cfg::kStagingBytes);
  uint64_t* full_bar = bars;                      // [kStages]  TMA -> MMA        (CG=2: the leader's is used)
  uint64_t* empty_bar = bars + kStages;           // [kStages]  MMA -> TMA        (CG=2: commit multicast to both)
  uint64_t* tmem_full_bar = bars + 2 * kStages;   // [2]        MMA -> epilogue   (CG=2: commit multicast to both)
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;   // [2]        epilogue -> MMA   (CG=2: both CTAs arrive on leader)
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = (CG == 2) ? cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], CG);   // leader: arrive.expect_tx ; peer: plain arrive after issuing its loads
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full_bar[s], 1);
      mbar_init(&tmem_empty_bar[s], kNumEpiWarps * CG);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    if constexpr (CG == 2) tmem_alloc_2sm<Cfg::kTmemCols>(tmem_base_smem);
    else tmem_alloc<Cfg::kTmemCols>(tmem_base_smem);
  }
  __syncwarp();
  tcgen05_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_base_smem;

  UnitSched sched;
  sched.init(p, CG, static_cast<int>(gridDim.x), static_cast<int>(blockIdx.x));

  if (warp < 4) {
  // warpgroup 0: TMA producer, MMA issuer, TMEM allocator, one idle warp.  (With > 8 epilogue warps the register file is
  // re-partitioned per warpgroup: ptxas budgets the code dominated by a setmaxnreg with that value, so the instruction
  // sits at the top of each role branch and every warp of the warpgroup executes it.)
  if constexpr (kEpiRegs > 0) asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
  if (warp == 0 && lane == 0) {
    // ===================== TMA producer (every CTA) =====================
    int stage = 0;
    uint32_t phase = 0;
    while (sched.next()) {
      const int m0 = sched.m_tile() * TILE_M + cta_rank * BLOCK_M;
      // ragged last column tile: when <= BLOCK_N/2 columns remain the MMA runs at half width (see the issuer), and the
      // CTAs of a pair split THAT width (rows beyond N are zero-filled by TMA and never touch DRAM)
      const int nt0 = sched.n_tile() * BLOCK_N;
      const bool narrow = BLOCK_N == 256 && (p.N - nt0) <= BLOCK_N / 2;
      const int n0 = nt0 + cta_rank * (narrow ? Cfg::kBRows / 2 : Cfg::kBRows);
      for (int kb = sched.kb0; kb < sched.kb1; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem_tiles + stage * Cfg::kStageBytes;
        uint8_t* sb = sa + Cfg::kABytes;
        if (leader) mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes * CG);
        const int k0 = kb * BLOCK_K;
        auto load = [&](const CUtensorMap* m, void* dst, int c0, int c1) {
#if MDT_TMA_HINT
          if constexpr (CG == 2) tma_load_2d_2sm_hint(m, &full_bar[stage], dst, c0, c1, kL2EvictLast);
          else tma_load_2d_hint(m, &full_bar[stage], dst, c0, c1, kL2EvictLast);
#else
          if constexpr (CG == 2) tma_load_2d_2sm(m, &full_bar[stage], dst, c0, c1);
          else tma_load_2d(m, &full_bar[stage], dst, c0, c1);
#endif
        };
        if constexpr (!A_MN) {
          load(&tmap_a, sa, k0, m0);  // box {64 k, 128 rows}
        } else {
#pragma unroll
          for (int j = 0; j < BLOCK_M / 64; ++j)  // boxes {64 mn, 64 k}
            load(&tmap_a, sa + j * (64 * BLOCK_K * 2), m0 + j * 64, k0);
        }
        if constexpr (!B_MN) {
          load(&tmap_b, sb, k0, n0);  // box {64 k, BLOCK_N / CG rows}
        } else {
#pragma unroll
          for (int j = 0; j < Cfg::kBRows / 64; ++j)
            load(&tmap_b, sb + j * (64 * BLOCK_K * 2), n0 + j * 64, k0);
        }
        if constexpr (CG == 2) {
          if (!leader) mbar_arrive_cluster(&full_bar[stage], 0);
        }
        if (++stage == kStages) stage = 0, phase ^= 1;
      }
    }
  } else if (warp == 1 && lane == 0 && leader) {
    // ===================== MMA issuer (leader CTA only) =====================
    constexpr uint32_t idesc_full = make_idesc_bf16(TILE_M, BLOCK_N, A_MN ? 1 : 0, B_MN ? 1 : 0);
    constexpr uint32_t idesc_half = make_idesc_bf16(TILE_M, BLOCK_N / 2, A_MN ? 1 : 0, B_MN ? 1 : 0);
    int stage = 0;
    uint32_t phase = 0;
    int as = 0;
    uint32_t aphase = 0;
#ifdef MDT_GEMM_PROF  // MMA issuer of CTA 0: cycles waiting for the epilogue / for a full stage / issuing
    long long mt[3] = {0, 0, 0}, mt_prev = clock64();
#define MDT_MPROF(i) { const long long t_now = clock64(); mt[i] += t_now - mt_prev; mt_prev = t_now; }
#else
#define MDT_MPROF(i)
#endif
    while (sched.next()) {
      mbar_wait(&tmem_empty_bar[as], aphase ^ 1);
      tcgen05_fence_after();
      MDT_MPROF(0)  // accumulator stage released by the epilogue
      const uint32_t tmem_d = tmem_base + as * BLOCK_N;
      const bool narrow = BLOCK_N == 256 && (p.N - sched.n_tile() * BLOCK_N) <= BLOCK_N / 2;
      const uint32_t idesc = narrow ? idesc_half : idesc_full;  // half-width MMAs on a ragged last column tile
      for (int kb = sched.kb0; kb < sched.kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tcgen05_fence_after();
        MDT_MPROF(1)  // operands of this k-block landed
        const uint32_t sa = smem_u32(smem_tiles + stage * Cfg::kStageBytes);
        const uint32_t sb = sa + Cfg::kABytes;
#pragma unroll
        for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
          // K-major: +32 B per UMMA_K inside the 128B swizzle row; SBO = 8 rows * 128 B.
          // MN-major: +16 k-rows * 128 B per UMMA_K; LBO = one 64-wide MN atom (64 k-rows * 128 B), SBO = 8 k-rows.
          const uint64_t da = A_MN ? make_smem_desc_sw128(sa + k * (UMMA_K * 128), 64 * BLOCK_K * 2, 1024)
                                   : make_smem_desc_sw128(sa + k * (UMMA_K * 2), 16, 1024);
          const uint64_t db = B_MN ? make_smem_desc_sw128(sb + k * (UMMA_K * 128), 64 * BLOCK_K * 2, 1024)
                                   : make_smem_desc_sw128(sb + k * (UMMA_K * 2), 16, 1024);
          const uint32_t acc = (kb > sched.kb0 || k > 0) ? 1u : 0u;
          if constexpr (CG == 2) umma_bf16_2sm(tmem_d, da, db, idesc, acc);
          else umma_bf16(tmem_d, da, db, idesc, acc);
        }
        // frees the smem slot (in both CTAs) once these MMAs retire
        if constexpr (CG == 2) umma_commit_2sm(&empty_bar[stage]); else umma_commit(&empty_bar[stage]);
        if (++stage == kStages) stage = 0, phase ^= 1;
        MDT_MPROF(2)  // issue of the k-block's MMAs + commit
      }
      // accumulator complete -> epilogue warps (of both CTAs)
      if constexpr (CG == 2) umma_commit_2sm(&tmem_full_bar[as]); else umma_commit(&tmem_full_bar[as]);
      if (++as == 2) as = 0, aphase ^= 1;
    }
#ifdef MDT_GEMM_PROF
    if (blockIdx.x == 0) {
      g_gemm_prof[6] = static_cast<float>(mt[0]);  // [6] wait accumulator release, [7] wait full stage
      g_gemm_prof[7] = static_cast<float>(mt[1]);
      g_gemm_prof[8] = static_cast<float>(mt[2]);  // [8] issue
    }
#endif
#undef MDT_MPROF
  }
  } else {
    // ===================== epilogue (every CTA: its own 128 accumulator rows) =====================
    if constexpr (kEpiRegs > 0) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kEpiRegs));
    const uint32_t stg = smem_u32(staging) + (warp - 4) * kStgFloats * 4;
#define MDT_EPI_LOOP(E) \
  epilogue_loop<E, BLOCK_N, CG>(p, sched, tmem_full_bar, tmem_empty_bar, tmem_base, stg, warp, cta_rank, lane)
    switch (p.epi) {
      case EPI_STORE: MDT_EPI_LOOP(EPI_STORE); break;
      case EPI_GELU: MDT_EPI_LOOP(EPI_GELU); break;
      case EPI_GATE_RESID: MDT_EPI_LOOP(EPI_GATE_RESID); break;
      case EPI_DGELU: MDT_EPI_LOOP(EPI_DGELU); break;
      case EPI_ATOMIC: MDT_EPI_LOOP(EPI_ATOMIC); break;
      default: MDT_EPI_LOOP(99); break;  // debug: accumulators drained, nothing stored
    }
#undef MDT_EPI_LOOP
  }

  __syncwarp();  // role branches above are per-lane; reconverge before the .aligned barriers below
  tcgen05_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    if constexpr (CG == 2) tmem_dealloc_2sm<Cfg::kTmemCols>(tmem_base);
    else tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(f);
  });
  return fn;
}

// 2-D bf16 tensor map: dims {inner, outer}, row stride ld elements, box {box_inner, box_outer}, 128B swizzle.
static int make_tmap(CUtensorMap* m, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner,
                     uint32_t box_outer) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return MDT_ERR_DRIVER;
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? MDT_OK : MDT_ERR_TMAP;
}

int make_row_tile_tmap(void* m, const void* ptr, unsigned long long rows, unsigned long long row_elems,
                       unsigned box_cols, unsigned box_rows) {
  if ((box_cols != 64 && box_cols != 32) || box_rows > 256 || (row_elems % 8) || (reinterpret_cast<uintptr_t>(ptr) & 15))
    return MDT_ERR_ARG;
  if (box_cols == 64)
    return make_tmap(static_cast<CUtensorMap*>(m), ptr, row_elems, rows, row_elems, box_cols, box_rows);
  // 32 columns = 64-byte rows: SWIZZLE_64B
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return MDT_ERR_DRIVER;
  cuuint64_t dims[2] = {row_elems, rows};
  cuuint64_t strides[1] = {row_elems * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(static_cast<CUtensorMap*>(m), CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims,
                   strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? MDT_OK : MDT_ERR_TMAP;
}

int make_token_tile_tmap(void* m, const void* ptr, unsigned long long rows, unsigned long long row_elems,
                         unsigned box_chunks, unsigned box_row_blocks) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return MDT_ERR_DRIVER;
  if (rows % 8 || row_elems % 8 || (reinterpret_cast<uintptr_t>(ptr) & 15)) return MDT_ERR_ARG;
  cuuint64_t dims[4] = {8, 8, row_elems / 8, rows / 8};
  cuuint64_t strides[3] = {row_elems * 2, 16, 8 * row_elems * 2};
  cuuint32_t box[4] = {8, 8, box_chunks, box_row_blocks};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(static_cast<CUtensorMap*>(m), CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims,
                   strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? MDT_OK : MDT_ERR_TMAP;
}
extern int g_sm_budget;
static int g_num_sms = 0;
static int num_sms_device() {
  if (!g_num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = kNumSMsDefault;
  }
  return g_num_sms;
}
// SMs the persistent grid is sized for: the device's, or the budget set by mdt_set_sm_budget (rounded down to even so
// that SM pairs stay whole)
static int num_sms() {
  const int n = num_sms_device();
  if (g_sm_budget > 0 && g_sm_budget < n) return g_sm_budget >= 2 ? (g_sm_budget & ~1) : 2;
  return n;
}

extern int g_gemm_last_config, g_gemm_configs_seen;
extern int g_tile_order, g_split_rule;
static thread_local GemmPlan* g_plan_out = nullptr;  // set by gemm_plan() around a gemm_launch() call
template <int BLOCK_N, bool A_MN, bool B_MN, int CG>
static int launch(const mdt_gemm_args& a, cudaStream_t stream) {
  using Cfg = GemmCfg<BLOCK_N, CG>;
  constexpr int TILE_M = BLOCK_M * CG;
  GemmParams p;
  p.M = a.M, p.N = a.N, p.K = a.K;
  p.epi = a.epi, p.act = a.act;
  p.num_m_tiles = (a.M + TILE_M - 1) / TILE_M;
  p.num_n_tiles = (a.N + BLOCK_N - 1) / BLOCK_N;
  p.num_kb = (a.K + BLOCK_K - 1) / BLOCK_K;
  p.narrow_last = (BLOCK_N == 256 && p.num_n_tiles > 1 && (a.N - (p.num_n_tiles - 1) * BLOCK_N) <= BLOCK_N / 2) ? 1 : 0;
  p.out = a.out, p.ldo = a.ldo, p.out_fp32 = a.out_fp32;
  p.bias = a.bias;
  p.aux = a.aux, p.ld_aux = a.ld_aux;
  p.resid = a.resid, p.ld_resid = a.ld_resid;
  p.gate = a.gate, p.ld_gate = a.ld_gate, p.rows_per_group = a.rows_per_group > 0 ? a.rows_per_group : 1;
  p.colsum = a.epi == EPI_DGELU ? a.colsum : nullptr;
  const int tiles = p.num_m_tiles * p.num_n_tiles;
  const int groups = num_sms() / CG;  // CTA groups resident at once (1 CTA per SM)
  // k-slices: only for the accumulate epilogue (fp32 red.add).  Units = tiles x slices are dealt round robin to the
  // CTA groups, so a launch lasts ceil(units / groups) unit-times; a unit costs its k-blocks plus a fixed part (pipeline
  // fill, the exposed share of the tile's reduction epilogue: ~4 k-blocks' worth).  Pick the slice count that minimises
  // waves x (num_kb / slices + 4).  (The first rule took the smallest of {1,2,3,4,6,8,12,16,32} slices that filled
  // >= 90 % of the last wave: 3 slices for the 90-tile fc1 / fc2 wgrads = 270 units on 74 pairs = 4 waves of 171
  // k-blocks, where 4 slices give 5 waves of 128: -6 %; MDT_GEMM_SPLITS=r1 keeps that rule for A/B.)
  int splits = 1;
  if (a.epi == EPI_ATOMIC && g_split_rule == 1) {
    double best_eff = 0.0;
    const int cand[9] = {1, 2, 3, 4, 6, 8, 12, 16, 32};
    for (int c : cand) {
      if (c > p.num_kb) break;
      const long long u = static_cast<long long>(tiles) * c;
      const double eff = static_cast<double>(u) / (static_cast<double>((u + groups - 1) / groups) * groups);
      if (eff > best_eff + 0.03) best_eff = eff, splits = c;
      if (eff >= 0.9) break;
    }
  } else if (a.epi == EPI_ATOMIC) {
    double best = 0.0;
    for (int c = 1; c <= 32 && c <= p.num_kb; ++c) {
      const long long u = static_cast<long long>(tiles) * c;
      const double t = static_cast<double>((u + groups - 1) / groups) * (static_cast<double>(p.num_kb) / c + 4.0);
      if (c == 1 || t < best * 0.995) best = t, splits = c;  // ties and near-ties go to fewer slices (fewer reductions)
    }
  }
  p.streamk = splits;
  // tile order of the half-width last column: paired (locality, default for the non-accumulating epilogues) or LPT
  p.pair_halves = (p.narrow_last && a.epi != EPI_ATOMIC && g_tile_order != 1) ? 1 : 0;
  const long long units = gemm_units_per_slice(p) * splits;
  const int grid = static_cast<int>(units < groups ? units : groups) * CG;
  if (g_plan_out) {  // mdt_gemm_plan: report the decisions, launch nothing (no device, no driver needed)
    *g_plan_out = GemmPlan{BLOCK_N, CG, splits, p.pair_halves, p.narrow_last, p.num_m_tiles, p.num_n_tiles, p.num_kb,
                           units, grid};
    return MDT_OK;
  }
  if (grid <= 0) return MDT_OK;
  CUtensorMap ta, tb;
  int rc;
  // A: K-major stored [M, K] ld=lda -> dims {K, M}, box {64, 128}; MN-major stored [K, M] -> dims {M, K}, box {64, 64}
  rc = A_MN ? make_tmap(&ta, a.A, a.M, a.K, a.lda, 64, 64) : make_tmap(&ta, a.A, a.K, a.M, a.lda, 64, BLOCK_M);
  if (rc) return rc;
  rc = B_MN ? make_tmap(&tb, a.B, a.N, a.K, a.ldb, 64, 64) : make_tmap(&tb, a.B, a.K, a.N, a.ldb, 64, Cfg::kBRows);
  if (rc) return rc;
  auto kern = gemm_tcgen05_kernel<BLOCK_N, A_MN, B_MN, CG>;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes) != cudaSuccess)
      return MDT_ERR_CUDA;
    attr_set = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid), cfg.blockDim = dim3(kNumThreads), cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr, cfg.numAttrs = 1;
  if (cudaLaunchKernelEx(&cfg, kern, ta, tb, p) != cudaSuccess) return MDT_ERR_CUDA;
  g_gemm_last_config = BLOCK_N * 10 + CG;
  g_gemm_configs_seen |= 1 << ((BLOCK_N / 64 - 2) * 2 + (CG - 1));
  return cudaGetLastError() == cudaSuccess ? MDT_OK : MDT_ERR_CUDA;
}

int g_gemm_last_config = 0;  // BLOCK_N * 10 + CG of the last launch (tests assert the 2-CTA instances ran)
int g_gemm_configs_seen = 0;  // bit (BLOCK_N/64 - 2) * 2 + (CG - 1) per instance launched since the last reset
static int g_force_cg = 0;  // 0 = auto, 1 / 2 = forced (MDT_GEMM_CG env, for A/B measurements)
int g_tile_order = 0;  // half-width column tiles: 0 = paired m-major units (default), 1 = LPT (MDT_GEMM_ORDER=lpt, for A/B)
int g_split_rule = 0;  // k-slices of the accumulating GEMMs: 0 = wave-time model (default), 1 = round-1 rule (MDT_GEMM_SPLITS=r1)

template <bool A_MN, bool B_MN>
static int dispatch_n(const mdt_gemm_args& a, cudaStream_t stream) {
  static bool env_read = false;
  if (!env_read) {
    const char* e = getenv("MDT_GEMM_CG");
    if (e) g_force_cg = atoi(e);
    const char* o = getenv("MDT_GEMM_ORDER");
    if (o && o[0] == 'l') g_tile_order = 1;
    const char* sr = getenv("MDT_GEMM_SPLITS");
    if (sr && sr[0] == 'r') g_split_rule = 1;
    env_read = true;
  }
  // SM pairs (256-row tiles) whenever there are at least two 128-row panels; single CTAs for skinny problems
  int cg = (a.M > BLOCK_M) ? 2 : 1;
  if (g_force_cg == 1 || g_force_cg == 2) cg = g_force_cg;
  // Tile width: 256 columns unless the problem is narrower.  Measured on B200 (tools/probe_gemm3.py): for N = 1152
  // a 256-wide tile (5 tiles, 10 % padding) beats 192 (no padding) and 128 by 1.2-1.3x, because the epilogue and
  // the per-tile pipeline fill are amortised over twice the MMA work.
  int best = 256;
  if (a.N <= 128) best = 128;
  else if (a.N <= 192 && !(cg == 2 && B_MN)) best = 192;
  if (a.block_n == 128 || a.block_n == 192 || a.block_n == 256) best = a.block_n;
  if (cg == 2 && B_MN && best == 192) best = 128;
  if (cg == 2) {
    switch (best) {
      case 256: return launch<256, A_MN, B_MN, 2>(a, stream);
      case 192:
        if constexpr (!B_MN) return launch<192, A_MN, B_MN, 2>(a, stream);
        return MDT_ERR_ARG;
      default: return launch<128, A_MN, B_MN, 2>(a, stream);
    }
  }
  switch (best) {
    case 256: return launch<256, A_MN, B_MN, 1>(a, stream);
    case 192: return launch<192, A_MN, B_MN, 1>(a, stream);
    default: return launch<128, A_MN, B_MN, 1>(a, stream);
  }
}

int gemm_launch(const mdt_gemm_args& a, cudaStream_t stream) {
  if (a.M <= 0 || a.N <= 0 || a.K <= 0) return MDT_ERR_ARG;
  if ((a.lda % 8) || (a.ldb % 8)) return MDT_ERR_ARG;                                   // TMA: 16-byte row strides
  if ((reinterpret_cast<uintptr_t>(a.A) & 15) || (reinterpret_cast<uintptr_t>(a.B) & 15)) return MDT_ERR_ARG;
  if (a.epi == EPI_ATOMIC && !a.out_fp32) return MDT_ERR_ARG;
  if (a.epi == EPI_GATE_RESID && (!a.gate || !a.resid)) return MDT_ERR_ARG;
  if (a.epi == EPI_DGELU && !a.aux) return MDT_ERR_ARG;
  // vectorised epilogue accesses need 32-column chunks to start 16B-aligned
  if (a.ldo % 8) return MDT_ERR_ARG;
  if (a.a_mn && a.b_mn) return dispatch_n<true, true>(a, stream);
  if (!a.a_mn && a.b_mn) return dispatch_n<false, true>(a, stream);
  if (!a.a_mn && !a.b_mn) return dispatch_n<false, false>(a, stream);
  return MDT_ERR_ARG;  // (MN, K) is never needed by this path
}

int gemm_plan(const mdt_gemm_args& a, GemmPlan* out) {
  GemmPlan pl = {};
  g_plan_out = &pl;
  const int rc = gemm_launch(a, nullptr);
  g_plan_out = nullptr;
  if (rc == MDT_OK && out) *out = pl;
  return rc;
}

}  // namespace mdt

#ifdef MDT_GEMM_PROF
extern "C" int mdt_debug_gemm_prof(float* out12) {  // development build only: last launch's phase cycles
  return cudaMemcpyFromSymbol(out12, mdt::g_gemm_prof, 12 * sizeof(float)) == cudaSuccess ? 0 : -1;
}
#endif
