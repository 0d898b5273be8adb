#!/bin/bash
# Artifacts for profiles/: ncu launch list of one training step, ncu --set full of the dominant GEMM, full bench line.
mkdir -p gpurun_out
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_train256.csv python tools/profile_step.py 256 32 > gpurun_out/ncu_launch.log 2>&1
echo "launch list exit $?"
python tools/summarize_launches.py gpurun_out/launches_train256.csv > gpurun_out/launches_train256.md 2>/dev/null; head -12 gpurun_out/launches_train256.md
# dominant kernel: fc1 forward GEMM <256,0,0,2> (-s skips the conditioning GEMMs and the first blocks)
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 30 -c 4 \
    -o gpurun_out/prof_gemm_r01 python tools/profile_step.py 256 32 > gpurun_out/ncu_full.log 2>&1
echo "full capture exit $?"
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_train256.json 2> gpurun_out/bench_train256.err
echo "bench exit $?"; cat gpurun_out/bench_train256.json | cut -c1-2500
timeout 600 python bench.py --gpus 1 --workload sampler --steps 10 --warmup 3 > gpurun_out/bench_sampler.json 2> gpurun_out/bench_sampler.err
echo "sampler bench exit $?"; cat gpurun_out/bench_sampler.json | cut -c1-1500; tail -3 gpurun_out/bench_sampler.err
