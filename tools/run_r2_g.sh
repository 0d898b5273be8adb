#!/bin/bash
# round 2, GPU job G (8 GPUs): the gradient all-reduce alone under NCCL settings, then two full bench lines at N=8
N=${1:-8}
probe() { env "$@" timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29660 tools/allreduce_probe.py 2>/dev/null | grep "^N="; }
probe A=1
probe NCCL_ALGO=Ring
probe NCCL_ALGO=NVLS
probe NCCL_ALGO=Tree
probe NCCL_MIN_NCHANNELS=32
probe NCCL_NVLS_ENABLE=0
