# final validation of the committed state: the driver's end-of-round sequence (whole GPU suite in ONE process, smoke, both bench arms)
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s9f_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/s9f_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s9f_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/s9f_smoke.log
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/s9f_bench_ref.log 2>&1
( time timeout 900 python bench.py ) > gpurun_out/s9f_bench.log 2>&1
tail -3 gpurun_out/s9f_tests.log; tail -2 gpurun_out/s9f_smoke.log; tail -c 300 gpurun_out/s9f_bench_ref.log; tail -c 200 gpurun_out/s9f_bench.log
