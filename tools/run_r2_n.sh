#!/bin/bash
# A/B on ONE box: the tree of commit 2327c5f (job F: 122.1 ms) vs the current tree
one() { ( cd $1 && timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-sub 2>/dev/null | python -c "
import json,sys; d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); print('$2', round(d['value'],1), 'samples/s', round(d['ms_per_step'],2), 'ms gemm frac', round(d['roofline']['frac'],3), d['clocks']['sm_mhz'])" ); }
one _ab/old "old tree"
one . "new tree"
one _ab/old "old tree"
one . "new tree"
