set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r2b_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2b_tests.log
python tools/profile_step.py 32768 72 2 > gpurun_out/r2b_prof.json 2> gpurun_out/r2b_prof.err
for v in 0 2; do
  RRTMGX_LW_GN=$v python tools/profile_step.py 32768 72 2 > gpurun_out/r2b_prof_lw$v.json 2> gpurun_out/r2b_prof_lw$v.err
done
python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r2b_bench.log 2>&1
