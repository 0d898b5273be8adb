// Probe (development tool, not part of the library): how fast can every SM fill its shared memory from L2 with TMA, and
// does cluster multicast raise that rate?  Models the GEMM's operand stream without the MMAs: persistent CTAs, a ring
// of 32 KB stages, per iteration each CTA receives a 16 KB "A" tile (always private) and a 16 KB "B" tile that is
//   mode 0: loaded privately by every CTA (what gemm_tcgen05_kernel does today);
//   mode 1: shared by the CS CTAs of a cluster: CTA r loads rows [r*128/CS, (r+1)*128/CS) and multicasts them to all.
// Output: bytes received per clock per SM and the chip-wide rate.   nvcc -arch=sm_100a -o tma_fill_probe tma_fill_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../maskdit_b200/csrc/common.cuh"
using namespace mdt;

constexpr int kStages = 5, kTile = 128 * 64 * 2;  // 16 KB: 128 rows x 64 bf16 (one 128B-swizzle panel)

__device__ __forceinline__ void tma_load_2d_mc(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(dst)), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
  uint32_t addr;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(addr) : "r"(smem_u32(bar)), "r"(cta));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(addr) : "memory");
}

template <int CS>
__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap ta, const __grid_constant__ CUtensorMap tb, int mode, int iters,
             int rows_total, long long* cycles_out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + kStages * 2 * kTile);
  uint64_t* empty = full + kStages;
  const uint32_t rank = CS > 1 ? cluster_ctarank() : 0u;
  const int cluster_id = blockIdx.x / CS;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], (mode == 1) ? CS : 1);  // a multicast slot is free when every CTA of the cluster consumed it
    }
    fence_barrier_init();
  }
  if (CS > 1) cluster_sync_all(); else __syncthreads();
  long long t0 = clock64();
  if (threadIdx.x == 0) {  // producer
    int stage = 0;
    uint32_t phase = 0;
    for (int it = 0; it < iters; ++it) {
      mbar_wait(&empty[stage], phase ^ 1);
      mbar_arrive_expect_tx(&full[stage], 2 * kTile);
      uint8_t* sa = smem + stage * 2 * kTile;
      uint8_t* sb = sa + kTile;
      const int k0 = (it * 64) % 4096;
      const int ra = ((blockIdx.x * 131 + it) * 128) % rows_total;      // private A rows
      const int rb = ((cluster_id * 977 + it * 7) * 128) % rows_total;  // B rows of this cluster
      tma_load_2d(&ta, &full[stage], sa, k0, ra);
      if (mode == 0 || CS == 1) {
        tma_load_2d(&ta, &full[stage], sb, k0, rb);
      } else {
        constexpr int part = 128 / CS;
        tma_load_2d_mc(&tb, &full[stage], sb + rank * part * 128, k0, rb + rank * part, static_cast<uint16_t>((1u << CS) - 1));
      }
      if (++stage == kStages) stage = 0, phase ^= 1;
    }
  } else if (threadIdx.x == 32) {  // consumer: waits for the stage, touches nothing, releases it
    int stage = 0;
    uint32_t phase = 0;
    for (int it = 0; it < iters; ++it) {
      mbar_wait(&full[stage], phase);
      if (mode == 1 && CS > 1) {
        for (uint32_t c = 0; c < CS; ++c) mbar_arrive_remote(&empty[stage], c);
      } else {
        mbar_arrive(&empty[stage]);
      }
      if (++stage == kStages) stage = 0, phase ^= 1;
    }
  }
  __syncthreads();
  if (CS > 1) cluster_sync_all();
  if (threadIdx.x == 0) cycles_out[blockIdx.x] = clock64() - t0;
}

typedef CUresult (*PFN_enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                            const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                            CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static CUtensorMap make_map(PFN_enc enc, void* ptr, uint64_t cols, uint64_t rows, uint32_t box_rows) {
  CUtensorMap m;
  cuuint64_t dims[2] = {cols, rows}, strides[1] = {cols * 2};
  cuuint32_t box[2] = {64, box_rows}, es[2] = {1, 1};
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
  return m;
}

template <int CS>
static void run(PFN_enc enc, void* buf, int rows, int mode, int iters, int sms) {
  CUtensorMap ta = make_map(enc, buf, 4096, rows, 128), tb = make_map(enc, buf, 4096, rows, 128 / CS);
  const int smem = kStages * 2 * kTile + 1024 + 256;
  auto kern = probe_kernel<CS>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  int grid = sms / CS * CS;
  if (CS > 1) {  // how many clusters can be co-resident?
    cudaLaunchConfig_t q = {};
    q.gridDim = dim3(grid), q.blockDim = dim3(128), q.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension, at[0].val.clusterDim.x = CS, at[0].val.clusterDim.y = 1, at[0].val.clusterDim.z = 1;
    q.attrs = at, q.numAttrs = 1;
    int nclusters = 0;
    cudaOccupancyMaxActiveClusters(&nclusters, kern, &q);
    if (nclusters * CS < grid) grid = nclusters * CS;
  }
  long long* cyc;
  cudaMalloc(&cyc, grid * sizeof(long long));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid), cfg.blockDim = dim3(128), cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension, at[0].val.clusterDim.x = CS, at[0].val.clusterDim.y = 1, at[0].val.clusterDim.z = 1;
  cfg.attrs = at, cfg.numAttrs = 1;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0);
    cudaError_t err = cudaLaunchKernelEx(&cfg, kern, ta, tb, mode, iters, rows, cyc);
    cudaEventRecord(e1);
    if (err != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) { printf("CS=%d mode=%d launch/exec failed: %s\n", CS, mode, cudaGetErrorString(cudaGetLastError())); return; }
  }
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  long long* h = (long long*)malloc(grid * sizeof(long long));
  cudaMemcpy(h, cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost);
  double mean = 0;
  for (int i = 0; i < grid; ++i) mean += h[i];
  mean /= grid;
  const double bytes_per_cta = 2.0 * kTile * iters;
  printf("cluster %d  %-34s grid %3d CTAs: %6.1f B/clk received per SM, %6.2f TB/s received chip-wide (%.3f ms)\n", CS,
         mode == 0 ? "private A + private B (unicast)" : "private A + multicast shared B", grid, bytes_per_cta / mean,
         bytes_per_cta * grid / (ms * 1e-3) / 1e12, ms);
  cudaFree(cyc);
  free(h);
}

int main() {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  void* f = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q);
  PFN_enc enc = (PFN_enc)f;
  const int rows = 8192;  // 8192 x 4096 bf16 = 64 MB: L2-resident working set like the GEMM's operand panels
  void* buf;
  cudaMalloc(&buf, (size_t)rows * 4096 * 2);
  cudaMemset(buf, 0, (size_t)rows * 4096 * 2);
  const int iters = 4000;
  run<1>(enc, buf, rows, 0, iters, sms);
  run<2>(enc, buf, rows, 0, iters, sms);
  run<2>(enc, buf, rows, 1, iters, sms);
  run<4>(enc, buf, rows, 0, iters, sms);
  run<4>(enc, buf, rows, 1, iters, sms);
  run<8>(enc, buf, rows, 1, iters, sms);
  return 0;
}
