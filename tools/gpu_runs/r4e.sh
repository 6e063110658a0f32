# end-of-round validation of the committed state (the driver's sequence): GPU suite, smoke, both bench arms
python -m pytest tests -m gpu -x -q > gpurun_out/r4e_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r4e_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r4e_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r4e_smoke.log
( time python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/r4e_bench_ref.log 2>&1
( time python bench.py ) > gpurun_out/r4e_bench.log 2>&1
tail -3 gpurun_out/r4e_tests.log; tail -2 gpurun_out/r4e_smoke.log; tail -c 400 gpurun_out/r4e_bench_ref.log; tail -c 700 gpurun_out/r4e_bench.log
