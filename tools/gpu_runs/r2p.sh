# LW upward sweep streams the exp/tfn entries back from scratch instead of repeating the table look-up
python -m pytest tests/test_lw_gpu.py tests/test_glue_gpu.py -m gpu -x -q > gpurun_out/r2p_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2p_tests.log
python tools/profile_step.py 32768 72 2 > gpurun_out/r2p_prof.json 2> gpurun_out/r2p_prof.err
python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2p_bench.log 2>&1
