#!/bin/bash
# round 2, GPU job B: previously failing tests + the C++ step driver parity + host-overhead A/B (C driver vs Python engine)
timeout 900 python -m pytest tests -m gpu -q -s -k "attention_fwd_bwd or eval_cfg or state_dict_resume or xl2_eval or from_moments or ablation or grad_accum or c_driver or optimizer_state or entrypoints or train_from_lmdb or train_then" 2>&1 | grep -v "^$" | tail -n 120
for b in 256 128 64; do
for e in c py; do
  MDT_ENGINE=$e timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --batch-per-gpu $b 2>/dev/null | python -c "
import json,sys; d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); print('train256 B=$b engine=$e', round(d['value'],1), 'samples/s', round(d['ms_per_step'],2), 'ms  e2e', round(d['e2e']['value'],1), 'launches', d['gpu_launches'], 'gemm frac', round(d['roofline']['frac'],3), 'clk', d['clocks']['sm_mhz'])"
done; done
MDT_TRAIN_GRAPH=1 timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --batch-per-gpu 128 2>/dev/null | python -c "
import json,sys; d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); print('train256 B=128 graph', round(d['value'],1), 'samples/s', round(d['ms_per_step'],2), 'ms')"
