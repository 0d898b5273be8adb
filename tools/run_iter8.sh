#!/bin/bash
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "gemm" 2>&1 | tail -n 3
timeout 200 python tools/probe_gemm2.py 2>&1 | grep -E "CG=|rror"
timeout 300 python tools/probe_gemm.py time 2>&1 | grep TIME
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value'],1), 'samples/s', round(d['ms_per_step'],2), 'ms', 'gemm', round(d['roofline']['achieved']), 'share', round(d['roofline']['share_of_step'],3), 'clk', d['clocks']['sm_mhz'])"
