#!/bin/bash
mkdir -p gpurun_out
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_train512.csv python tools/profile_step.py 128 64 > gpurun_out/ncu_launch512.log 2>&1
echo "launch list exit $?"
python tools/summarize_launches.py gpurun_out/launches_train512.csv 2>/dev/null | head -22
