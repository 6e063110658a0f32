# McICA lanes = subcolumns of one column; SW aerosol-free cells skip the identity delta scaling
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r2g_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2g_tests.log
python tools/profile_step.py 32768 72 2 > gpurun_out/r2g_prof.json 2> gpurun_out/r2g_prof.err
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r2g_bench.log 2>&1
