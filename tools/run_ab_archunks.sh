#!/bin/bash
# A/B of the chunked all-reduce || AdamW pipeline on N GPUs of one box
N=${1:-2}
mkdir -p gpurun_out
for c in 8 1; do
MDT_AR_CHUNKS=$c timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); print('chunks=$c', round(d['value'],1), 'samples/s', round(d['ms_per_step'],2), 'ms', 'clk', d['clocks']['sm_mhz'])"
done
