#!/bin/bash
# Build libsegb200.so in-tree for sm_100a (nvcc cross-compiles without a GPU).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/../libsegb200.so"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --expt-relaxed-constexpr --extended-lambda -Xcompiler -fPIC ${SEGB_NVCC_EXTRA}"
OBJS=""
PIDS=""
for f in api dp fixedvar fixedvar_gibbs kmeans kmeans_mma fixedvar_mma fixedvar_filter score_fused frozen host_init diagnostics; do
  if [ -f "$HERE/$f.cu" ]; then
    stale=0
    for dep in "$HERE/$f.cu" "$HERE"/*.cuh "$HERE/../../include/segb200.h"; do
      if [ ! -f "$HERE/$f.o" ] || [ "$dep" -nt "$HERE/$f.o" ]; then stale=1; fi
    done
    if [ $stale = 1 ]; then
      rm -f "$HERE/$f.o"          # a failed compile must not leave a stale object to link
      $NVCC $FLAGS -c "$HERE/$f.cu" -o "$HERE/$f.o" &
      PIDS="$PIDS $!"
    fi
    OBJS="$OBJS $HERE/$f.o"
  fi
done
for pid in $PIDS; do
  wait $pid || { echo "build.sh: a compile job failed" >&2; exit 1; }
done
$NVCC -shared -gencode arch=compute_100a,code=sm_100a -o "$OUT" $OBJS -lcudart
echo "built $OUT"
