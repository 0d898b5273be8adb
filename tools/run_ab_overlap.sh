#!/bin/bash
for ov in 0 1 0 1; do
  MDT_OVERLAP=$ov timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('overlap=$ov', round(d['value'],1), 'samples/s', round(d['ms_per_step'],2), 'ms', 'gemm', round(d['roofline']['achieved']), 'clk', d['clocks']['sm_mhz'])"
done
