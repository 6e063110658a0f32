# real*4 host arrays (tests + informational e2e), full default bench line with cpu_baseline
python -m pytest tests -m gpu -x -q > gpurun_out/r2l_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2l_tests.log
( time python bench.py ) > gpurun_out/r2l_bench.log 2>&1
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2l_smoke.log 2>&1
