#!/bin/bash
# Build libsegb200.so in-tree for sm_100a (nvcc cross-compiles without a GPU).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/../libsegb200.so"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --expt-relaxed-constexpr --extended-lambda -Xcompiler -fPIC ${SEGB_NVCC_EXTRA}"
OBJS=""
for f in api dp fixedvar kmeans kmeans_mma fixedvar_mma; do
  if [ -f "$HERE/$f.cu" ]; then
    if [ ! -f "$HERE/$f.o" ] || [ "$HERE/$f.cu" -nt "$HERE/$f.o" ] || [ "$HERE/common.cuh" -nt "$HERE/$f.o" ] || [ "$HERE/mma_common.cuh" -nt "$HERE/$f.o" ] || [ "$HERE/../../include/segb200.h" -nt "$HERE/$f.o" ]; then
      $NVCC $FLAGS -c "$HERE/$f.cu" -o "$HERE/$f.o" &
    fi
    OBJS="$OBJS $HERE/$f.o"
  fi
done
wait
$NVCC -shared -gencode arch=compute_100a,code=sm_100a -o "$OUT" $OBJS -lcudart
echo "built $OUT"
