# the device KISS generator on its own (rrtmgx_debug_kiss) against the reference-source draws, the recorded real*4 range, the jump-ahead
timeout 300 python -m pytest tests/test_kiss_gpu.py -q > gpurun_out/s9b_kiss.log 2>&1; echo "tests rc=$?" >> gpurun_out/s9b_kiss.log
tail -25 gpurun_out/s9b_kiss.log
