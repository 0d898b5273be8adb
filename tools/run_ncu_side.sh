#!/bin/bash
# DRAM traffic / throughput and tensor-pipe activity of EVERY kernel of one training step (metrics subset, 1 step)
mkdir -p gpurun_out
ncu --profile-from-start off --clock-control none --csv \
    --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed \
    --log-file gpurun_out/side_kernels.csv python tools/profile_step.py 256 32 > gpurun_out/ncu_side.log 2>&1
echo "exit $?"; wc -l gpurun_out/side_kernels.csv; tail -2 gpurun_out/ncu_side.log
