#!/bin/bash
# usage: tools/build_variant.sh NAME "-DFLAG=1 ..."   -> build_variants/libdrt_cuda_NAME.so (for DRT_CUDA_LIB=... A/B runs on the GPU box)
# Builds the CUDA library with extra nvcc flags in a scratch object directory; the default build is left untouched.
set -e
NAME=$1; EXTRA=$2
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OUT=$ROOT/build_variants; OBJ=$OUT/obj_$NAME
mkdir -p $OBJ
NVFLAGS="$EXTRA -O3 -std=c++17 -lineinfo -prec-div=false -prec-sqrt=false -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -I$ROOT/include -I$ROOT/daily-ray-trace_b200/csrc"
pids=()
for f in $ROOT/daily-ray-trace_b200/csrc/*.cu; do
  nvcc $NVFLAGS -c $f -o $OBJ/$(basename ${f%.cu}).o & pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
nvcc $NVFLAGS -shared $OBJ/*.o -o $OUT/libdrt_cuda_$NAME.so -lcudart
echo built $OUT/libdrt_cuda_$NAME.so
