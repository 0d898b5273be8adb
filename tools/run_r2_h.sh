#!/bin/bash
# round 2, GPU job H (1 GPU): full GPU suite on the current tree, smoke(), per-shape GEMM table (GELU' stored by the
# forward epilogue), default bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -n 12
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -n 3
MDT_ENGINE=py timeout 600 python tools/gemm_shapes_step.py 256 32 2>&1 | head -24
timeout 900 python bench.py > gpurun_out/r02_bench_default_v2.json 2> gpurun_out/r02_bench_default_v2.err
echo "bench exit $?"; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_default_v2.json') if l.startswith('{')][-1])
print('C2', round(d['value'],1), 'samples/s', round(d['ms_per_step'],2), 'ms; e2e', round(d['e2e']['value'],1), '; gemm frac', round(d['roofline']['frac'],3), 'step frac', round(d['roofline']['step_frac'],3), 'clk', d['clocks'])
for k,v in d.get('sub',{}).items(): print(k, round(v['value'],1), v['unit'], round(v['ms_per_step'],2), 'ms')
print('cpu', d.get('cpu_baseline',{}).get('value'))
PY
