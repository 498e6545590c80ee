#!/bin/bash
# Build a variant of the CUDA library with extra nvcc flags into build_variants/<name>/libgdsp_b200.so
# (A/B experiments on the GPU box: GDSP_LIB_PATH=build_variants/<name>/libgdsp_b200.so python scripts/stage_bench.py ...)
#   scripts/build_variant.sh stage_deep -DGDSP_STAGE_DEEP
set -e
NAME=$1; shift
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OUT=$ROOT/build_variants/$NAME
mkdir -p $OUT/obj
cd $ROOT/genodsp_b200/csrc
for f in gdsp_*.cu; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 -Xcompiler -fPIC -I../../include "$@" -c $f -o $OUT/obj/${f%.cu}.o &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $OUT/libgdsp_b200.so $OUT/obj/*.o -cudart static
echo built $OUT/libgdsp_b200.so
