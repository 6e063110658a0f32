# sorted columns (by layer-1 pressure within the cloudy / cloud-free groups) against the caller's order
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r2c_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2c_tests.log
python tools/profile_step.py 32768 72 2 > gpurun_out/r2c_prof_sort.json 2> gpurun_out/r2c_prof_sort.err
RRTMGX_SORT=0 python tools/profile_step.py 32768 72 2 > gpurun_out/r2c_prof_nosort.json 2> gpurun_out/r2c_prof_nosort.err
for v in 0 2; do
  RRTMGX_LW_GN=$v python tools/profile_step.py 32768 72 2 > gpurun_out/r2c_prof_sort_lw$v.json 2> /dev/null
done
python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r2c_bench.log 2>&1
