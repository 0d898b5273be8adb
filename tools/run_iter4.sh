#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "gemm" 2>&1 | tail -n 3
timeout 200 python tools/probe_gemm2.py 2>&1 | grep -E "CG=|rror" | grep bn256
timeout 300 python tools/probe_gemm.py time 2>&1 | grep TIME
