# the two multi-GPU tests (skipped on one-GPU boxes) on two B200s, final code
timeout 600 python -m pytest tests/test_multigpu_gpu.py -q > gpurun_out/s9g_multigpu.log 2>&1; echo "tests rc=$?" >> gpurun_out/s9g_multigpu.log
tail -5 gpurun_out/s9g_multigpu.log
