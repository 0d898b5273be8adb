#!/bin/bash
# round 2, GPU job V (1 GPU): A/B of the half-width tile order (LPT vs paired m-major units) on one box:
# GEMM parity tests, per-shape in-step timing, step time, ncu --set full of four forward GEMMs with the paired order.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "gemm" 2>&1 | tail -n 5
for o in lpt paired lpt paired; do
  echo "== order [$o]"
  MDT_GEMM_ORDER=$o timeout 600 python bench.py --steps 12 --warmup 4 --no-cpu-baseline --no-sub 2>/dev/null | python -c "
import json,sys; d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); print('step [$o]', round(d['value'],1), 'samples/s', round(d['ms_per_step'],2), 'ms, gemm frac', round(d['roofline']['frac'],3), 'share', round(d['roofline']['share_of_step'],3), 'clk', d['clocks']['sm_mhz'])"
done
for o in lpt paired; do
  echo "== per-shape [$o]"
  MDT_ENGINE=py MDT_GEMM_ORDER=$o timeout 600 python tools/gemm_shapes_step.py 256 32 2>&1 | head -24
done
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 30 -c 4 \
    -o gpurun_out/prof_gemm_r02_paired -f python tools/profile_step.py 256 32 > gpurun_out/ncu_full_paired.log 2>&1
echo "full capture exit $?"
