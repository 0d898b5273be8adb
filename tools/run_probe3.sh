#!/bin/bash
for cg in 1 2; do MDT_GEMM_CG=$cg timeout 300 python tools/probe_gemm3.py 2>&1 | grep -E "CG=|rror|---"; done
