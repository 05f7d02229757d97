#!/bin/bash
# build_variant.sh NAME "EXTRA NVCC FLAGS": an experimental build of libfrz.so into build/variants/NAME.so (kernel tuning;
# selected at run time with FRZ_LIBRARY, see profiles/time_kernel.py)
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OUT=$ROOT/build/variants/$1
mkdir -p $OUT
cd $ROOT/free_range_zoo_b200/csrc
FLAGS="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC -Xptxas -v -I../../include $2"
for f in frz_api frz_wildfire frz_cyber frz_rideshare; do
  if [ "$f" = "${3:-frz_wildfire}" ] || [ ! -f $OUT/$f.o ]; then nvcc $FLAGS -c $f.cu -o $OUT/$f.o 2> $OUT/$f.ptxas.log & fi
done
wait
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o $ROOT/build/variants/$1.so $OUT/*.o
grep -A2 "${4:-wildfire_step_kernelILi16ELi7ELi0ELb0}" $OUT/${3:-frz_wildfire}.ptxas.log | grep -E "registers|spill" | head -4
