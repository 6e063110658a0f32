# validation of the committed state (the driver's end-of-round sequence) + the profiler evidence of the final kernels:
# whole GPU suite in ONE process, smoke, both bench arms, then (after the plain run exited 0) the ncu launch list of one
# bench step and one --set full capture with source of the top kernels
python -m pytest tests -m gpu -x -q > gpurun_out/s9a_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/s9a_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s9a_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/s9a_smoke.log
( time python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/s9a_bench_ref.log 2>&1
( time python bench.py ) > gpurun_out/s9a_bench.log 2>&1
CMD="python bench.py --steps 1 --warmup 1 --ncol 65536 --no-e2e --no-cpu --verify-cols 0"
$CMD > gpurun_out/s9a_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__sass_thread_inst_executed_ops_dadd_dmul_dfma_pred_on.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active \
    --clock-control none --csv --log-file gpurun_out/s9a_launches.csv $CMD > gpurun_out/s9a_ncu.log 2>&1
CMD2="python bench.py --steps 1 --warmup 1 --ncol 32768 --no-e2e --no-cpu --verify-cols 0"
ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:(sw_band_kernel<\(int\)17,)|(lw_band_kernel<\(int\)3,)|(mcica_kernel<rrtmgx::SwOptics)' -c 3 \
    -f -o gpurun_out/s9a_top $CMD2 > gpurun_out/s9a_ncu_full.log 2>&1
tail -3 gpurun_out/s9a_tests.log; tail -2 gpurun_out/s9a_smoke.log; tail -c 400 gpurun_out/s9a_bench_ref.log; tail -c 300 gpurun_out/s9a_bench.log; ls -la gpurun_out/s9a_*
