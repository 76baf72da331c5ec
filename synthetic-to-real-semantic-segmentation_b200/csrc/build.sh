#!/bin/bash
# Builds libs2r_b200.so (sm_100a only) next to the package.  Usage: build.sh [extra nvcc flags]
set -e
cd "$(dirname "$0")"
OUT=../libs2r_b200.so
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall --expt-relaxed-constexpr --extended-lambda"
mkdir -p build
pids=()
for f in *.cu; do
  o=build/${f%.cu}.o
  if [ ! -f "$o" ] || [ "$f" -nt "$o" ] || [ common.cuh -nt "$o" ] || [ conv_common.cuh -nt "$o" ] || [ dw_common.cuh -nt "$o" ] || [ tma_util.cuh -nt "$o" ] || [ bn_tail.cuh -nt "$o" ] || [ lerp.cuh -nt "$o" ] || [ ../../include/s2r_b200.h -nt "$o" ]; then
    $NVCC $FLAGS "$@" -c "$f" -o "$o" &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o $OUT build/*.o -lcudart_static -ldl -lrt -lpthread
echo "built $OUT"
