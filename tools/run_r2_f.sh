#!/bin/bash
# round 2, GPU job F (4 GPUs): DP equivalence (2 ranks) + gradient-exchange modes at N=4
timeout 600 python -m pytest tests/test_multigpu.py -q -s 2>&1 | grep -E "gradient rel-L2|DP_EQUIV|passed|failed" | tail -n 12
run() { lbl=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29650 bench.py --gpus 4 --steps 12 --warmup 4 --no-cpu-baseline --no-sub 2>/dev/null | python -c "
import json,sys; d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); print('N=4 [$lbl]', round(d['value'],1), 'samples/s', round(d['ms_per_step'],2), 'ms | ', d['config']['grad_allreduce'][:60])"
}
timeout 300 python bench.py --gpus 1 --steps 12 --warmup 4 --no-cpu-baseline --no-sub 2>/dev/null | python -c "
import json,sys; d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); print('N=1', round(d['value'],1), 'samples/s', round(d['ms_per_step'],2), 'ms')"
run "fp32 flat" MDT_GRAD_AR=fp32
run "bf16 flat" MDT_GRAD_AR=bf16
run "fp32 4 chunks pipelined with the optimizer" MDT_GRAD_AR=fp32 MDT_AR_CHUNKS=4
run "bf16 4 chunks pipelined with the optimizer" MDT_GRAD_AR=bf16 MDT_AR_CHUNKS=4
run "bf16 8 chunks pipelined with the optimizer" MDT_GRAD_AR=bf16 MDT_AR_CHUNKS=8
