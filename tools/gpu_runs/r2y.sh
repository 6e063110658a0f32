# McICA thresholds and SW cloud coefficients tiled by 32 columns
python -m pytest tests -m gpu -x -q > gpurun_out/r2y_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2y_tests.log
python tools/profile_step.py 65536 72 2 > gpurun_out/r2y_prof.json 2> gpurun_out/r2y_prof.err
python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2y_bench.log 2>&1
