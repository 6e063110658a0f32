# validation of the committed state (the driver's end-of-round sequence): whole GPU suite in ONE process, smoke, both bench
# arms; then, after the plain run of the same command exited 0, the ncu launch list (times only: one pass per kernel)
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t1c_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/t1c_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/t1c_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/t1c_smoke.log
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/t1c_bench_ref.log 2>&1
( time timeout 900 python bench.py ) > gpurun_out/t1c_bench.log 2>&1
CMD="python bench.py --steps 1 --warmup 1 --ncol 65536 --no-e2e --no-cpu --verify-cols 0"
timeout 300 $CMD > gpurun_out/t1c_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/t1c_launches.csv $CMD > gpurun_out/t1c_ncu.log 2>&1
tail -3 gpurun_out/t1c_tests.log; tail -2 gpurun_out/t1c_smoke.log; tail -c 400 gpurun_out/t1c_bench_ref.log; tail -c 300 gpurun_out/t1c_bench.log; ls -la gpurun_out/t1c_*
