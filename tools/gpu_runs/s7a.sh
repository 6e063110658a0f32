# validation of the committed state, the driver's sequence: whole GPU suite in ONE process, smoke, both bench arms
python -m pytest tests -m gpu -x -q > gpurun_out/s7a_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/s7a_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s7a_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/s7a_smoke.log
( time python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/s7a_bench_ref.log 2>&1
( time python bench.py ) > gpurun_out/s7a_bench.log 2>&1
tail -3 gpurun_out/s7a_tests.log; tail -2 gpurun_out/s7a_smoke.log; tail -c 500 gpurun_out/s7a_bench_ref.log; tail -c 300 gpurun_out/s7a_bench.log
