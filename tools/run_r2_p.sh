#!/bin/bash
# A/B on ONE box: round 1's final tree (fd5a030) vs the current tree, default bench line without sub-records
one() { ( cd $1 && timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline $3 2>/dev/null | python -c "
import json,sys; d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); r=d['roofline']; print('$2', round(d['value'],1), 'samples/s', round(d['ms_per_step'],2), 'ms | gemm frac', round(r['frac'],3), 'share', round(r['share_of_step'],3), d['clocks']['sm_mhz'])" ); }
one _ab/r01 "round-1 tree" ""
one . "round-2 tree" "--no-sub"
one _ab/r01 "round-1 tree" ""
one . "round-2 tree" "--no-sub"
