#!/bin/bash
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -x -k "gemm" 2>&1 | tail -n 3
timeout 900 python -m pytest tests/test_model_gpu.py -q -x 2>&1 | tail -n 3
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('train', round(d['value'],1), 'samples/s', round(d['ms_per_step'],2), 'ms', 'gemm', round(d['roofline']['achieved']), 'clk', d['clocks']['sm_mhz'])"
