#!/bin/bash
timeout 1200 python -m pytest tests -q -x -m gpu 2>&1 | tail -n 3
python tools/run_attn_time.py 2>&1 | head -3
python tools/timeline_step.py 256 32 2>&1 | grep -v -i warn | sed -n 1,5p
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('train', round(d['value'],1), 'samples/s', round(d['ms_per_step'],2), 'ms', 'gemm', round(d['roofline']['achieved']), 'clk', d['clocks']['sm_mhz'])"
MDT_LIB_PATH=$PWD/maskdit_b200/libmaskdit_b200_prev.so timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('train(prev lib, new python)', round(d['value'],1), 'samples/s', round(d['ms_per_step'],2), 'ms', 'gemm', round(d['roofline']['achieved']), 'clk', d['clocks']['sm_mhz'])"
