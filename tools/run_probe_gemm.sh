#!/bin/bash
# one process per group so that a trap in one group does not poison the others
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
for g in kk kmn mnmn epi time; do
  echo "=== $g ==="
  timeout 300 python tools/probe_gemm.py $g > gpurun_out/probe_gemm_$g.log 2>&1
  echo "exit $?"
  tail -n 40 gpurun_out/probe_gemm_$g.log
done
