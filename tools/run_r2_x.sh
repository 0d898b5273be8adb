#!/bin/bash
# round 2, GPU job X (1 GPU): the final tree (paired tile order) - full GPU suite, smoke(), default bench line,
# reference arm, ncu launch list of one training step.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -n 6
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -n 2
timeout 600 python bench.py > gpurun_out/r02_bench_final2.json 2> gpurun_out/r02_bench_final2.err; echo "bench exit $?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_arm2.json 2>/dev/null; echo "reference arm exit $?"
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/r02_bench_final2.json') if l.startswith('{')][-1])
print('C2', round(d['value'], 1), 'samples/s', round(d['ms_per_step'], 2), 'ms; e2e', round(d['e2e']['value'], 1), '; gemm frac', round(d['roofline']['frac'], 3), 'step frac', round(d['roofline']['step_frac'], 3), d['clocks'], 'launches', d['gpu_launches'], 'traffic', d['roofline']['traffic'], d['roofline']['traffic_source'])
for k, v in d.get('sub', {}).items():
    if 'error' in v:
        print('   ', k, v); continue
    print('   ', k, round(v['value'], 1), v['unit'], round(v['ms_per_step'], 2), 'ms', round(v['roofline'].get('step_frac', v['roofline'].get('frac', 0)), 3))
cb = d.get('cpu_baseline', {})
print('    cpu', cb.get('kind'), cb.get('value'), cb.get('cores'), {k: round(v['value'], 3) for k, v in cb.items() if isinstance(v, dict)})
r = json.loads([l for l in open('gpurun_out/r02_bench_reference_arm2.json') if l.startswith('{')][-1])
print('reference arm', r['cpu_baseline']['kind'], round(r['value'], 3), 'samples/s', r['cpu_baseline']['cores'], 'threads')
PY
timeout 240 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/r02_launches_train256_v2.csv python tools/profile_step.py 256 32 > gpurun_out/ncu_launch2.log 2>&1
echo "launch list exit $?"
python tools/summarize_launches.py gpurun_out/r02_launches_train256_v2.csv > gpurun_out/r02_launches_train256_v2.md 2>/dev/null; head -24 gpurun_out/r02_launches_train256_v2.md
