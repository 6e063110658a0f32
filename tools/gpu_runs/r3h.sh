# end-of-round validation (the driver's sequence) and the ncu evidence of the final kernels
python -m pytest tests -m gpu -x -q > gpurun_out/r3h_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r3h_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r3h_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r3h_smoke.log
( time python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/r3h_bench_ref.log 2>&1
( time python bench.py ) > gpurun_out/r3h_bench.log 2>&1
CMD="python bench.py --steps 1 --warmup 1 --ncol 65536 --no-e2e --no-cpu"
$CMD > gpurun_out/r3h_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread \
    --clock-control none --csv --log-file gpurun_out/r3h_launches.csv $CMD > gpurun_out/r3h_ncu.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:(sw_band_kernel<\(int\)(17|20),)|(lw_band_kernel<\(int\)(3|9),)|(mcica_kernel<rrtmgx::SwOptics)' -c 5 \
    -f -o gpurun_out/r3h_top $CMD > gpurun_out/r3h_ncu_full.log 2>&1
ls -la gpurun_out/ | tail -4
