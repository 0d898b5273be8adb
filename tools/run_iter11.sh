#!/bin/bash
# A/B of library builds within one box: GEMM parity tests, per-shape in-step GEMM timing, bench
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "gemm" 2>&1 | tail -n 3
for v in new nopf prev; do
  if [ $v = new ]; then unset MDT_LIB_PATH; else export MDT_LIB_PATH=$PWD/maskdit_b200/libmaskdit_b200_$v.so; fi
  echo "=== $v"
  python tools/gemm_shapes_step.py 256 32 2>&1 | head -28
  timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v', round(d['value'],1), 'samples/s', round(d['ms_per_step'],2), 'ms', 'gemm', round(d['roofline']['achieved']), 'share', round(d['roofline']['share_of_step'],3), 'clk', d['clocks']['sm_mhz'])"
done
