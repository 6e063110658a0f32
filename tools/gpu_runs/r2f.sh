# glue parity tests + full suite + bench after the clean-up of the prefetch experiment
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r2f_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2f_tests.log
python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r2f_bench.log 2>&1
