#!/bin/bash
mkdir -p gpurun_out
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_train256.csv python tools/profile_step.py 256 32 > gpurun_out/ncu_launch.log 2>&1
echo "launch list exit $?"
python tools/summarize_launches.py gpurun_out/launches_train256.csv 2>/dev/null | head -32
