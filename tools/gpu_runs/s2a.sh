# source-level ncu capture of the two dominant band kernels (stall reasons per line) on the round-2 entry state
CMD="python bench.py --steps 1 --warmup 1 --ncol 32768 --no-e2e --no-cpu"
$CMD > gpurun_out/s2a_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:(sw_band_kernel<\(int\)17,)|(lw_band_kernel<\(int\)3,)' -c 2 \
    -f -o gpurun_out/s2a_top $CMD > gpurun_out/s2a_ncu_full.log 2>&1
ls -la gpurun_out/ | tail -4; tail -3 gpurun_out/s2a_ncu_full.log
