#!/bin/bash
# A/B build of the GEMM with other compile-time settings:
#   tools/build_gemm_variant.sh 12                       -> maskdit_b200/libmaskdit_b200_e12.so    (-DMDT_EPI_WARPS=12)
#   tools/build_gemm_variant.sh cs -DMDT_EPI_CS=1        -> maskdit_b200/libmaskdit_b200_cs.so
# (select at run time with MDT_LIB_PATH=maskdit_b200/libmaskdit_b200_<suffix>.so)
set -e
S=$1; shift
if [ $# -eq 0 ]; then DEFS="-DMDT_EPI_WARPS=$S"; S=e$S; else DEFS="$@"; fi
cd "$(dirname "$0")/.."
python -m maskdit_b200.build > /dev/null
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --expt-relaxed-constexpr -Xcompiler -fPIC \
     $DEFS -c maskdit_b200/csrc/gemm_tcgen05.cu -o maskdit_b200/build/gemm_tcgen05_$S.o
objs=$(ls maskdit_b200/build/*.o | grep -v "gemm_tcgen05" )
nvcc -shared -o maskdit_b200/libmaskdit_b200_$S.so $objs maskdit_b200/build/gemm_tcgen05_$S.o -Xcompiler -fPIC
echo built maskdit_b200/libmaskdit_b200_$S.so
