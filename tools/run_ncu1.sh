#!/bin/bash
mkdir -p gpurun_out
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_train256.csv python tools/profile_step.py 256 32 > gpurun_out/ncu_launch.log 2>&1
echo "launch list exit $?"; tail -n 3 gpurun_out/ncu_launch.log
python tools/summarize_launches.py gpurun_out/launches_train256.csv | head -40
# full capture of the dominant kernel: 3 forward-shape GEMM launches inside the profiled step
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 40 -c 4 \
    -o gpurun_out/prof_gemm python tools/profile_step.py 256 32 > gpurun_out/ncu_full.log 2>&1
echo "full capture exit $?"; tail -n 3 gpurun_out/ncu_full.log; ls -la gpurun_out/*.ncu-rep
