// extern "C" surface that is not tied to one kernel file: status strings, ABI version, GEMM entry.
#include "gemm.h"

namespace mdt {
extern int g_gemm_last_config, g_gemm_configs_seen;
int g_sm_budget = 0;  // mdt_set_sm_budget: SMs the persistent kernels may occupy (0 = all)
}

extern "C" {

const char* mdt_status_string(int status) {
  switch (status) {
    case MDT_OK: return "ok";
    case MDT_ERR_ARG: return "invalid argument (shape / alignment / null pointer)";
    case MDT_ERR_CUDA: return "CUDA launch failure";
    case MDT_ERR_DRIVER: return "cuTensorMapEncodeTiled driver entry point unavailable";
    case MDT_ERR_TMAP: return "tensor map encode failed";
    case MDT_ERR_UNSUPPORTED: return "unsupported configuration";
    default: return "unknown status";
  }
}

int mdt_abi_version(void) { return 2; }  // 2: mdt_gemm_args.colsum, step driver, gradient exchange

int mdt_set_sm_budget(int n) {
  if (n < 0) return MDT_ERR_ARG;
  mdt::g_sm_budget = n;
  return MDT_OK;
}
int mdt_get_sm_budget(void) { return mdt::g_sm_budget; }

int mdt_gemm_last_config(void) { return mdt::g_gemm_last_config; }
int mdt_gemm_configs_seen(int reset) {
  const int v = mdt::g_gemm_configs_seen;
  if (reset) mdt::g_gemm_configs_seen = 0;
  return v;
}

int mdt_gemm_bf16(const mdt_gemm_args* args, void* stream) {
  if (!args || !args->A || !args->B || !args->out) return MDT_ERR_ARG;
  return mdt::gemm_launch(*args, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
