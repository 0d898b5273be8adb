#!/bin/bash
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "gemm" 2>&1 | tail -n 2
timeout 900 python bench.py --gpus 1 --workload train512 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_train512.json 2> gpurun_out/bench_train512.err
echo "exit $?"; cat gpurun_out/bench_train512.json | cut -c1-1800; tail -n 5 gpurun_out/bench_train512.err
nvidia-smi --query-gpu=memory.used --format=csv
