# new bench.py (verification, roofline from counters, native real*4 glue arm) + the ncu launch list the counters come from
( time python bench.py --steps 3 --warmup 3 ) > gpurun_out/s4a_bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/s4a_bench.log
CMD="python bench.py --steps 1 --warmup 1 --ncol 65536 --no-e2e --no-cpu --verify-cols 0"
$CMD > gpurun_out/s4a_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__sass_thread_inst_executed_ops_dadd_dmul_dfma_pred_on.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active \
    --clock-control none --csv --log-file gpurun_out/s4a_launches.csv $CMD > gpurun_out/s4a_ncu.log 2>&1
tail -c 600 gpurun_out/s4a_bench.log; ls -la gpurun_out/s4a_launches.csv
