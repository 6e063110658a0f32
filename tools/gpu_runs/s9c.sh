# ncu --set full with source of the final top kernels (SW band 17, LW band 3, SW McICA) on a SMALL launch (16 384 columns:
# kernel replay saves and restores the state the kernel writes, which is what made s9a's launch list crawl)
CMD="python bench.py --steps 1 --warmup 1 --ncol 16384 --no-e2e --no-cpu --verify-cols 0"
timeout 120 $CMD > gpurun_out/s9c_plain.log 2>&1 || exit 1
timeout 420 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:(sw_band_kernel<\(int\)17,)|(lw_band_kernel<\(int\)3,)|(mcica_kernel<rrtmgx::SwOptics)' -c 3 \
    -f -o gpurun_out/s9c_top $CMD > gpurun_out/s9c_ncu_full.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/s9c_*; tail -3 gpurun_out/s9c_ncu_full.log
