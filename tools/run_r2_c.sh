#!/bin/bash
# round 2, GPU job C (2 GPUs): remaining parity tests, DP equivalence, exchange-mode A/B at N=2
timeout 900 python -m pytest tests -m gpu -q -s -k "eval_cfg or xl2_ or c_driver or multigpu or train_step_cuda_graph" 2>&1 | grep -v "^$" | tail -n 60
run() { # $1 = label, rest = env
  lbl=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29650 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-sub 2>/dev/null | python -c "
import json,sys; d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); print('N=2 [$lbl]', round(d['value'],1), 'samples/s', round(d['ms_per_step'],2), 'ms | ', d['config']['grad_allreduce'])"
}
timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-sub 2>/dev/null | python -c "
import json,sys; d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); print('N=1', round(d['value'],1), 'samples/s', round(d['ms_per_step'],2), 'ms')"
run "torch fp32 flat" MDT_COLLECTIVE=torch
run "mdt fp32 flat" MDT_COLLECTIVE=mdt
run "mdt bf16 flat" MDT_GRAD_AR=bf16
run "mdt fp32 overlap 8 ctas" MDT_OVERLAP=1 MDT_COMM_CTAS=8
run "mdt bf16 overlap 8 ctas" MDT_OVERLAP=1 MDT_COMM_CTAS=8 MDT_GRAD_AR=bf16
run "mdt bf16 overlap 4 ctas" MDT_OVERLAP=1 MDT_COMM_CTAS=4 MDT_GRAD_AR=bf16
run "mdt bf16 overlap 16 ctas" MDT_OVERLAP=1 MDT_COMM_CTAS=16 MDT_GRAD_AR=bf16
