#!/bin/bash
# the roofline measurement with library-side event pairs: repeatability on one box
for i in 1 2 3; do
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-sub 2>/dev/null | python -c "
import json,sys; d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); r=d['roofline']; print('run $i', round(d['value'],1), 'samples/s', round(d['ms_per_step'],2), 'ms | gemm', round(r['achieved'],1), 'TF/s frac', round(r['frac'],3), 'share', round(r['share_of_step'],3), 'launches', r['launches_per_step'], d['clocks']['sm_mhz'])"
done
