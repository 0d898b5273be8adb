#!/bin/bash
# round 2, GPU job Y (1 GPU): k-slice rule of the accumulating (wgrad) GEMMs - wave-time model (new default) vs the
# round-1 rule (MDT_GEMM_SPLITS=r1) on one box, then the full GPU suite on the new default.
mkdir -p gpurun_out
for r in r1 model r1 model; do
  MDT_GEMM_SPLITS=$r timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-sub 2>/dev/null | python -c "
import json,sys; d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); print('step [$r]', round(d['value'],1), 'samples/s', round(d['ms_per_step'],2), 'ms, gemm frac', round(d['roofline']['frac'],3), 'share', round(d['roofline']['share_of_step'],3), 'clk', d['clocks']['sm_mhz'], 'loss', round(d['config']['final_loss'],4))"
done
for r in r1 model; do
  echo "== per-shape [$r]"
  MDT_ENGINE=py MDT_GEMM_SPLITS=$r timeout 300 python tools/gemm_shapes_step.py 256 32 2>&1 | grep -E "GEMM total|atomic"
done
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -n 4
