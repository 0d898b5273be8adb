#!/bin/bash
# round 2, GPU job J (2 GPUs): full GPU suite incl. the 2-rank equivalence test, 2-rank train.py / generate.py smoke,
# default bench line at N=1 and N=2 (sub-records under torchrun)
mkdir -p gpurun_out /tmp/mdt_j
timeout 1800 python -m pytest tests -q -s -m gpu 2>&1 | grep -E "passed|failed|VAE decode rel-L2|8-bit image|worst grad|FAILED|rel-L2 vs 1-GPU" | tail -n 30
cat > /tmp/mdt_j/cfg.yaml <<'Y'
data: {dataset: imagenet256-latent, category: lmdb, resolution: 16, num_channels: 4, root: none, feat_path: None}
model: {precond: edm, model_type: DiT-S/2, in_size: 16, in_channels: 4, num_classes: 1000, use_decoder: True,
        ext_feature_dim: 0, pad_cls_token: False, mask_ratio: 0.5, mask_ratio_fn: constant, mask_ratio_min: 0,
        mae_loss_coef: 0.1, class_dropout_prob: 0.1}
train: {tf32: False, amp: True, batchsize: 8, grad_accum: 1, epochs: 1, lr: 0.0001, lr_rampup_kimg: 0, xflip: False,
        max_num_steps: 4}
log: {log_every: 2, ckpt_every: 4, tag: t}
Y
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29670 train.py --config /tmp/mdt_j/cfg.yaml --synthetic --max_steps 4 --results_dir /tmp/mdt_j/res 2>&1 | grep -E "Train Loss|Error|error" | tail -n 4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29671 generate.py --config /tmp/mdt_j/cfg.yaml --ckpt_path /tmp/mdt_j/res/checkpoints/0000004.pt --seeds 0-9 --num_steps 4 --max_batch_size 4 --results_dir /tmp/mdt_j/samples 2>&1 | grep -E "wrote|Error|error" | tail -n 4
ls /tmp/mdt_j/samples | wc -l
timeout 900 python bench.py > gpurun_out/r02_bench_n1_final.json 2> gpurun_out/r02_bench_n1_final.err; echo "bench N=1 exit $?"
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29672 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_bench_n2_final.json 2> gpurun_out/r02_bench_n2_final.err; echo "bench N=2 exit $?"
python - <<'PY'
import json
for f in ('gpurun_out/r02_bench_n1_final.json', 'gpurun_out/r02_bench_n2_final.json'):
    try:
        d = json.loads([l for l in open(f) if l.startswith('{')][-1])
    except Exception as e:
        print(f, 'NO JSON', e); continue
    print(f, 'N', d['n_gpus'], round(d['value'], 1), 'samples/s', round(d['ms_per_step'], 2), 'ms; e2e', round(d['e2e']['value'], 1), '; gemm frac', round(d['roofline']['frac'], 3), 'step frac', round(d['roofline']['step_frac'], 3), d['clocks'], '|', d['config']['grad_allreduce'][:70])
    for k, v in d.get('sub', {}).items():
        print('   ', k, round(v['value'], 1), v['unit'], round(v['ms_per_step'], 2), 'ms')
    print('    cpu', d.get('cpu_baseline', {}).get('value'), {k: v for k, v in d.get('cpu_baseline', {}).items() if k.startswith('c')})
PY
tail -n 5 gpurun_out/r02_bench_n2_final.err
