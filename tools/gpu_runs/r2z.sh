# end-of-round validation: the driver's sequence (tests, smoke, reference arm, bench)
python -m pytest tests -m gpu -x -q > gpurun_out/r2z_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2z_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2z_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2z_smoke.log
( time python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/r2z_bench_ref.log 2>&1
( time python bench.py ) > gpurun_out/r2z_bench.log 2>&1
