#!/bin/bash
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -x -k "attention or gate_bwd" 2>&1 | tail -n 3
bash tools/run_iter6.sh 2>&1 | grep attn
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value'],1), 'samples/s', round(d['ms_per_step'],2), 'ms', 'gemm', round(d['roofline']['achieved']), 'share', round(d['roofline']['share_of_step'],3), 'clk', d['clocks']['sm_mhz'])"
