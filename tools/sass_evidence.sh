#!/bin/bash
# SASS evidence for north_star items 2, 3, 5: which kernels of librrtmgx.so hold bulk-copy (TMA, non-tensor form:
# UBLKCP), mbarrier (SYNCS) and warp-shuffle (SHFL) instructions, with the first occurrences in context.
# Usage: tools/sass_evidence.sh > profiles/s8_sass_bulk_copy_and_shuffles.txt
SO="$(dirname "$0")/../geosradiation_gridcomp_b200/librrtmgx.so"
SASS=$(mktemp)
cuobjdump -sass "$SO" > "$SASS"
echo "# cuobjdump -sass geosradiation_gridcomp_b200/librrtmgx.so (sm_100a), $(date -u +%F)"
echo "# instruction counts per kernel (kernels without any of the three are not listed)"
awk '/Function :/{fn=$3} /UBLKCP/{u[fn]++} /SYNCS/{s[fn]++} /SHFL/{h[fn]++} END{for(f in u)k[f]=1; for(f in s)k[f]=1; for(f in h)k[f]=1; for(f in k) printf "%-100s UBLKCP %3d  SYNCS %3d  SHFL %3d\n", f, u[f], s[f], h[f]}' "$SASS" | grep rrtmgx | c++filt | sort
for fn in _ZN6rrtmgx14sw_down_kernelILi17ELi4EEEvNS_10SwBandArgsE _ZN6rrtmgx14lw_band_kernelILi3ELi2ELi80ELi4EEEvNS_10LwBandArgsE; do
  echo; echo "## $(echo $fn | c++filt)"
  awk -v fn="$fn" '/Function :/{on=($3==fn)} on' "$SASS" | grep -n -E "UBLKCP|SYNCS|SHFL" | head -12
done
rm -f "$SASS"
