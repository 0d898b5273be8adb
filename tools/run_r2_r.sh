#!/bin/bash
# default bench line (with sub-records) at 4 GPUs
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29672 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r02_bench_n4.json 2> gpurun_out/r02_bench_n4.err; echo "bench N=4 exit $?"
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/r02_bench_n4.json') if l.startswith('{')][-1])
print('N', d['n_gpus'], round(d['value'], 1), 'samples/s', round(d['ms_per_step'], 2), 'ms; e2e', round(d['e2e']['value'], 1), '; gemm frac', round(d['roofline']['frac'], 3), d['clocks'], '|', d['config']['grad_allreduce'][:90])
for k, v in d.get('sub', {}).items():
    print('   ', k, round(v['value'], 1), v['unit'], round(v['ms_per_step'], 2), 'ms')
PY
