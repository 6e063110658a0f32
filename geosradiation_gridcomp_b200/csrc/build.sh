#!/bin/bash
# Builds geosradiation_gridcomp_b200/librrtmgx.so for sm_100a (no other target, no CPU path).
# --fmad=false: every expression that feeds a Fortran int() truncation must be the same IEEE
# sequence as the reference (see common.cuh).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/../librrtmgx.so"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo --fmad=false -Xcompiler -fPIC -Xcompiler -O2 $RRTMGX_NVCC_FLAGS"
mkdir -p "$HERE/build"
SRCS="api.cu lw.cu kiss_jump.cpp tables.cpp"
DEFS=""
if [ -f "$HERE/sw.cu" ]; then SRCS="$SRCS sw.cu"; DEFS="-DRRTMGX_WITH_SW"; fi
pids=()
for f in $SRCS; do
  o="$HERE/build/${f%.*}.o"
  if [ ! -f "$o" ] || [ "$HERE/$f" -nt "$o" ] || [ -n "$(find "$HERE" -maxdepth 1 \( -name '*.h' -o -name '*.cuh' \) -newer "$o")" ] || [ "$HERE/../../include/rrtmgx.h" -nt "$o" ]; then
    ( $NVCC $FLAGS $DEFS -x cu -c "$HERE/$f" -o "$o" ) &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait $p; done
OBJS=""
for f in $SRCS; do OBJS="$OBJS $HERE/build/${f%.*}.o"; done
$NVCC -shared -gencode arch=compute_100a,code=sm_100a -o "$OUT" $OBJS -ldl
echo "built $OUT"
