#!/bin/bash
mkdir -p gpurun_out
ncu --set full --import-source on --clock-control none -k regex:attn_tc --launch-skip 4 --launch-count 4 \
    -o gpurun_out/prof_attn_r01 -f python tools/attn_shapes.py > gpurun_out/ncu_attn.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_attn.log; ls -la gpurun_out/*.ncu-rep
