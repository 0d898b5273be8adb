#!/bin/bash
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -x -k attention 2>&1 | tail -n 5
python tools/run_attn_time.py 2>&1 | head -3
