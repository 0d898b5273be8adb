#!/bin/bash
mkdir -p gpurun_out
export MDT_GEMM_CG=1
ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 2 -c 1 -o gpurun_out/prof_gemm_k64 python tools/one_gemm.py 64 > gpurun_out/ncu2a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 2 -c 1 -o gpurun_out/prof_gemm_k1152 python tools/one_gemm.py 1152 > gpurun_out/ncu2b.log 2>&1
tail -2 gpurun_out/ncu2a.log gpurun_out/ncu2b.log; ls -la gpurun_out/*.ncu-rep
