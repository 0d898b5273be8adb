#!/bin/bash
# 2-CTA GEMM probe: correctness groups then timing, each in its own process with a short timeout
mkdir -p gpurun_out
export MDT_GEMM_CG=2
for g in kk kmn mnmn epi time; do
  echo "=== CG=2 $g ==="
  timeout 120 python tools/probe_gemm.py $g > gpurun_out/probe_cg2_$g.log 2>&1
  echo "exit $?"
  grep -E "OK|BAD|TIME|rror|timeout" gpurun_out/probe_cg2_$g.log | head -n 30
done
export MDT_GEMM_CG=1
echo "=== CG=1 time ==="
timeout 120 python tools/probe_gemm.py time 2>&1 | grep TIME
