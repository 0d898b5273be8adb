#!/bin/bash
# A/B build of the GEMM with a different number of epilogue warps: tools/build_gemm_variant.sh 12 -> maskdit_b200/libmaskdit_b200_e12.so
# (select at run time with MDT_LIB_PATH=maskdit_b200/libmaskdit_b200_e12.so)
set -e
E=$1
cd "$(dirname "$0")/.."
python -m maskdit_b200.build > /dev/null
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --expt-relaxed-constexpr -Xcompiler -fPIC \
     -DMDT_EPI_WARPS=$E -c maskdit_b200/csrc/gemm_tcgen05.cu -o maskdit_b200/build/gemm_tcgen05_e$E.o
objs=$(ls maskdit_b200/build/*.o | grep -v "gemm_tcgen05" )
nvcc -shared -o maskdit_b200/libmaskdit_b200_e$E.so $objs maskdit_b200/build/gemm_tcgen05_e$E.o -Xcompiler -fPIC
echo built maskdit_b200/libmaskdit_b200_e$E.so
