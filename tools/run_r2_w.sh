#!/bin/bash
# round 2, GPU job W (1 GPU): L2-policy variants of the GEMM on one box (tools/build_gemm_variant.sh):
#   cs = epilogue traffic evict-first, hint = operand TMA loads evict-last, cshint = both.
mkdir -p gpurun_out
L=$PWD/maskdit_b200/libmaskdit_b200
MDT_LIB_PATH=${L}_cshint.so timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "gemm" 2>&1 | tail -n 3
for v in "" _cs _hint _cshint "" _cshint; do
  MDT_LIB_PATH=$L$v.so timeout 600 python bench.py --steps 12 --warmup 4 --no-cpu-baseline --no-sub 2>/dev/null | python -c "
import json,sys; d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); print('step [${v:-base}]', round(d['value'],1), 'samples/s', round(d['ms_per_step'],2), 'ms, gemm frac', round(d['roofline']['frac'],3), 'share', round(d['roofline']['share_of_step'],3), 'clk', d['clocks']['sm_mhz'])"
done
MDT_LIB_PATH=${L}_cshint.so timeout 600 ncu --profile-from-start off --clock-control none --csv \
    --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum -k regex:gemm_tcgen05 -s 30 -c 4 \
    --log-file gpurun_out/r02_gemm_dram_cshint.csv python tools/profile_step.py 256 32 > gpurun_out/ncu_cshint.log 2>&1
echo "ncu exit $?"; grep -v "^==" gpurun_out/r02_gemm_dram_cshint.csv | cut -d, -f5,12- | head -20
