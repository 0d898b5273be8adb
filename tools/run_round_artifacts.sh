#!/bin/bash
# Artifacts for profiles/: full GPU test-suite, ncu launch list of one training step, ncu --set full of the dominant
# GEMM, the default bench line, the sampler and 512-px bench lines.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -x -m gpu 2>&1 | tail -n 3
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_train256.csv python tools/profile_step.py 256 32 > gpurun_out/ncu_launch.log 2>&1
echo "launch list exit $?"
python tools/summarize_launches.py gpurun_out/launches_train256.csv > gpurun_out/launches_train256.md 2>/dev/null; head -14 gpurun_out/launches_train256.md
# dominant kernel: forward GEMM <256,0,0,2> (-s skips the conditioning GEMMs and the first blocks)
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 30 -c 4 \
    -o gpurun_out/prof_gemm_r01 -f python tools/profile_step.py 256 32 > gpurun_out/ncu_full.log 2>&1
echo "full capture exit $?"
timeout 900 python bench.py > gpurun_out/bench_train256.json 2> gpurun_out/bench_train256.err
echo "bench exit $?"; cat gpurun_out/bench_train256.json | cut -c1-2500
timeout 600 python bench.py --workload sampler --no-cpu-baseline > gpurun_out/bench_sampler.json 2> gpurun_out/bench_sampler.err
echo "sampler bench exit $?"; cat gpurun_out/bench_sampler.json | cut -c1-1200
timeout 900 python bench.py --workload train512 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_train512.json 2> gpurun_out/bench_train512.err
echo "train512 bench exit $?"; cat gpurun_out/bench_train512.json | cut -c1-600
