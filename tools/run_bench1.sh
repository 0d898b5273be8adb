#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py smoke 2>&1 | tail -3
echo "=== bench train256 ==="
timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/bench_train256.json 2> gpurun_out/bench_train256.err
echo "exit $?"; tail -c 3000 gpurun_out/bench_train256.json; tail -n 5 gpurun_out/bench_train256.err
echo "=== pytest -m gpu (single process, as the driver runs it) ==="
timeout 900 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -n 8
