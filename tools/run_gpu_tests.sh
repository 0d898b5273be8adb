#!/bin/bash
# Run each GPU kernel test group in its own process (a device trap in one group must not poison the others).
mkdir -p gpurun_out
: > gpurun_out/kernel_tests.log
for t in test_gemm_kk test_gemm_dgrad test_gemm_wgrad_streamk test_gemm_epilogues test_mask_indices_bit_exact \
         test_patch_embed_fwd_bwd test_timestep_freq test_pointwise test_ln_modulate_fwd_bwd test_gate_bwd \
         test_attention_fwd_bwd test_unmask_fwd_bwd test_edm_loss_and_grad test_heun_and_adamw; do
  echo "=== $t ===" | tee -a gpurun_out/kernel_tests.log
  timeout 300 python -m pytest tests/test_kernels_gpu.py -q -x -k "$t" 2>&1 | tail -n 25 | tee -a gpurun_out/kernel_tests.log
done
