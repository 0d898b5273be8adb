#!/bin/bash
# round 2, GPU job K (2 GPUs): overlapped exchange with TWO communicators (few-CTA one during the backward, full-width one
# for the gradients that become final at the end) vs the default chunk-pipelined exchange
run() { lbl=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29650 bench.py --gpus 2 --steps 12 --warmup 4 --no-cpu-baseline --no-sub 2>/dev/null | python -c "
import json,sys; d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); print('N=2 [$lbl]', round(d['value'],1), 'samples/s', round(d['ms_per_step'],2), 'ms')"
}
timeout 300 python bench.py --gpus 1 --steps 12 --warmup 4 --no-cpu-baseline --no-sub 2>/dev/null | python -c "
import json,sys; d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); print('N=1', round(d['value'],1), 'samples/s', round(d['ms_per_step'],2), 'ms')"
run "bf16, 4 chunks after the backward (default)" A=1
run "bf16 overlapped, 4 SMs reserved" MDT_OVERLAP=1 MDT_COMM_CTAS=4
run "bf16 overlapped, 8 SMs reserved" MDT_OVERLAP=1 MDT_COMM_CTAS=8
run "bf16 overlapped, 16 SMs reserved" MDT_OVERLAP=1 MDT_COMM_CTAS=16
timeout 600 python -m pytest tests/test_multigpu.py -q -s 2>&1 | grep -E "overlapped|DP_EQUIV|passed|failed" | cut -c1-110
