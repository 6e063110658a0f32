#!/bin/bash
# An alternative build of the library for A/B measurements: tools/build_variant.sh NAME "EXTRA NVCC FLAGS"
# -> geosradiation_gridcomp_b200/csrc/build_NAME/librrtmgx_NAME.so (git-ignored; swapped in by a tools/gpu_runs script).
set -e
NAME="$1"; EXTRA="$2"
HERE="$(cd "$(dirname "$0")/../geosradiation_gridcomp_b200/csrc" && pwd)"
B="$HERE/build_$NAME"; mkdir -p "$B"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo --fmad=false -Xcompiler -fPIC -Xcompiler -O2 -DRRTMGX_WITH_SW $EXTRA"
pids=()
for f in api.cu lw.cu kiss_jump.cpp tables.cpp sw.cu; do
  ( $NVCC $FLAGS -x cu -c "$HERE/$f" -o "$B/${f%.*}.o" 2> "$B/${f%.*}.log" ) & pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -shared -gencode arch=compute_100a,code=sm_100a -o "$B/librrtmgx_$NAME.so" "$B"/*.o -ldl
echo "built $B/librrtmgx_$NAME.so"
