# McICA sweep stops above the last layer that holds cloud
python -m pytest tests -m gpu -x -q > gpurun_out/r2r_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2r_tests.log
python tools/profile_step.py 32768 72 2 > gpurun_out/r2r_prof.json 2> gpurun_out/r2r_prof.err
python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2r_bench.log 2>&1
