#!/bin/bash
# same-box repeatability of the headline number (the pool's boxes gave 121.3 .. 126.1 ms for the same tree)
for i in 1 2 3; do
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-sub 2>/dev/null | python -c "
import json,sys; d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); print('run $i C engine', round(d['value'],1), 'samples/s', round(d['ms_per_step'],2), 'ms gemm frac', round(d['roofline']['frac'],3), d['clocks']['sm_mhz'])"
done
MDT_ENGINE=py timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-sub 2>/dev/null | python -c "
import json,sys; d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); print('py engine', round(d['value'],1), 'samples/s', round(d['ms_per_step'],2), 'ms gemm frac', round(d['roofline']['frac'],3), d['clocks']['sm_mhz'])"
nvidia-smi --query-gpu=name,power.limit,power.max_limit,clocks.max.sm,temperature.gpu --format=csv
