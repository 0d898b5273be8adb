#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_train256_n$N.json 2> gpurun_out/bench_train256_n$N.err
echo "exit $?"; tail -c 2500 gpurun_out/bench_train256_n$N.json; grep -v -E "^\s*$|Warning|warn" gpurun_out/bench_train256_n$N.err | tail -n 8
