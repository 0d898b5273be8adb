#!/bin/bash
# round 2, GPU job L (1 GPU): the final tree — build check, full GPU suite, smoke(), default bench line, reference arm
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu 2>&1 | tail -n 6
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -n 2
timeout 900 python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; echo "bench exit $?"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2>/dev/null; echo "reference arm exit $?"
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/r02_bench_final.json') if l.startswith('{')][-1])
print('C2', round(d['value'], 1), 'samples/s', round(d['ms_per_step'], 2), 'ms; e2e', round(d['e2e']['value'], 1), '; gemm frac', round(d['roofline']['frac'], 3), 'step frac', round(d['roofline']['step_frac'], 3), d['clocks'], 'launches', d['gpu_launches'])
for k, v in d.get('sub', {}).items():
    print('   ', k, round(v['value'], 1), v['unit'], round(v['ms_per_step'], 2), 'ms', round(v['roofline'].get('step_frac', v['roofline'].get('frac', 0)), 3))
cb = d.get('cpu_baseline', {})
print('    cpu', cb.get('kind'), cb.get('value'), cb.get('cores'), {k: round(v['value'], 3) for k, v in cb.items() if isinstance(v, dict)})
r = json.loads([l for l in open('gpurun_out/r02_bench_reference_arm.json') if l.startswith('{')][-1])
print('reference arm', r['cpu_baseline']['kind'], round(r['value'], 3), 'samples/s', r['cpu_baseline']['cores'], 'threads')
PY
