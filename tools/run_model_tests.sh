#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/model_tests.log
for t in test_train_loss_and_grads_vs_reference_golden test_generic_autograd_path_matches_fused \
         test_eval_cfg_and_sampler_vs_reference_golden test_xl2_config1_forward_vs_reference_golden \
         test_train_step_matches_oracle_adamw_and_ema; do
  echo "=== $t ===" | tee -a gpurun_out/model_tests.log
  timeout 600 python -m pytest tests/test_model_gpu.py -q -x -s -k "$t" 2>&1 | tail -n 40 | tee -a gpurun_out/model_tests.log
done
for t in test_patch_embed_fwd_bwd test_heun_and_adamw; do
  timeout 300 python -m pytest tests/test_kernels_gpu.py -q -x -k "$t" 2>&1 | tail -n 5 | tee -a gpurun_out/model_tests.log
done
