// Internal declarations shared by gemm_tcgen05.cu and api.cu.
#pragma once
#include <cuda_runtime.h>

#include "../../include/maskdit_b200.h"

namespace mdt {

enum { EPI_STORE = MDT_EPI_STORE, EPI_GELU = MDT_EPI_GELU, EPI_GATE_RESID = MDT_EPI_GATE_RESID,
       EPI_DGELU = MDT_EPI_DGELU, EPI_ATOMIC = MDT_EPI_ATOMIC };
enum { ACT_NONE = MDT_ACT_NONE, ACT_SILU = MDT_ACT_SILU };

struct GemmParams {
  int M, N, K;
  int epi, act;
  int num_m_tiles, num_n_tiles, num_kb;
  int streamk;
  void* out;
  int ldo, out_fp32;
  const float* bias;
  void* aux;
  int ld_aux;
  const float* resid;
  int ld_resid;
  const float* gate;
  int ld_gate, rows_per_group;
};

int gemm_launch(const mdt_gemm_args& a, cudaStream_t stream);

}  // namespace mdt
