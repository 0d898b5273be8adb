#!/bin/bash
# round 2, GPU job D (1 GPU): full GPU suite, r02 ncu evidence (launch list, --set full of the GEMM, per-kernel DRAM/tensor
# metrics), the default bench line with sub-records, background-optimizer A/B.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -n 25
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/r02_launches_train256.csv python tools/profile_step.py 256 32 > gpurun_out/ncu_launch.log 2>&1
echo "launch list exit $?"
python tools/summarize_launches.py gpurun_out/r02_launches_train256.csv > gpurun_out/r02_launches_train256.md 2>/dev/null; head -30 gpurun_out/r02_launches_train256.md
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 30 -c 4 \
    -o gpurun_out/prof_gemm_r02 -f python tools/profile_step.py 256 32 > gpurun_out/ncu_full.log 2>&1
echo "full capture exit $?"
ncu --profile-from-start off --clock-control none --csv \
    --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed \
    --log-file gpurun_out/r02_side_kernels.csv python tools/profile_step.py 256 32 > gpurun_out/ncu_side.log 2>&1
echo "side metrics exit $?"
timeout 900 python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err
echo "bench exit $?"; cut -c1-6000 gpurun_out/r02_bench_default.json
MDT_OVERLAP=1 MDT_COMM_CTAS=4 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-sub 2>/dev/null | python -c "
import json,sys; d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); print('N=1 background optimizer, 4 SMs reserved:', round(d['value'],1), 'samples/s', round(d['ms_per_step'],2), 'ms')"
MDT_OVERLAP=1 MDT_COMM_CTAS=2 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-sub 2>/dev/null | python -c "
import json,sys; d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); print('N=1 background optimizer, 2 SMs reserved:', round(d['value'],1), 'samples/s', round(d['ms_per_step'],2), 'ms')"
# GEMM epilogue-warp variants (8 default / 12 / 16): per-shape in-step timing + step time, same box
for v in "" _e12 _e16; do
  echo "== GEMM variant [${v:-e8}]"
  MDT_ENGINE=py MDT_LIB_PATH=$PWD/maskdit_b200/libmaskdit_b200$v.so timeout 600 python tools/gemm_shapes_step.py 256 32 2>&1 | head -34
  MDT_LIB_PATH=$PWD/maskdit_b200/libmaskdit_b200$v.so timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-sub 2>/dev/null | python -c "
import json,sys; d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); print('step [${v:-e8}]', round(d['value'],1), 'samples/s', round(d['ms_per_step'],2), 'ms, gemm frac', round(d['roofline']['frac'],3), 'clk', d['clocks']['sm_mhz'])"
done
