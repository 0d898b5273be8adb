#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -x -k "attention" 2>&1 | tail -n 15
timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_iter.json 2> gpurun_out/bench_iter.err
echo "bench exit $?"; python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/bench_iter.json'))
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['roofline'], d['clocks'], d['e2e'])
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/bench_iter.err').read()[-3000:])
PY
timeout 600 python -m pytest tests/test_model_gpu.py -q -x 2>&1 | tail -n 4
