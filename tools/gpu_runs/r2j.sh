# SW band kernels with lanes along the g-points of a column and the per-cell scratch [band][lay][plane][nc][ng]
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r2j_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2j_tests.log
for v in 0 1 2 3; do
  RRTMGX_SW_GN=$v python tools/profile_step.py 32768 72 2 > gpurun_out/r2j_prof_sw$v.json 2> gpurun_out/r2j_prof_sw$v.err
done
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r2j_bench.log 2>&1
