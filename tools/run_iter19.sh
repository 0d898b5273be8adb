#!/bin/bash
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k attention 2>&1 | tail -n 8
for p in 1 0; do echo "== sw=$p"; MDT_ATTN_SW=$p timeout 300 python tools/run_attn_time.py 2>&1 | head -3; done
