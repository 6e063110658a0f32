# SW split into an upward kernel and a streaming (bulk-copy ring) downward kernel: parity tests, then timing sweep
python -m pytest tests/test_sw_gpu.py tests/test_fullsize_gpu.py tests/test_glue_gpu.py -m gpu -x -q > gpurun_out/s2b_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/s2b_tests.log
tail -4 gpurun_out/s2b_tests.log
python tools/sweep.py 65536 72 "fused:RRTMGX_SW_SPLIT=0" "split_u0d0:RRTMGX_SW_SPLIT=1,RRTMGX_SW_UP=0,RRTMGX_SW_DOWN=0" \
   "split_u1d0:RRTMGX_SW_UP=1,RRTMGX_SW_DOWN=0" "split_u2d0:RRTMGX_SW_UP=2,RRTMGX_SW_DOWN=0" "split_u1d1:RRTMGX_SW_UP=1,RRTMGX_SW_DOWN=1" \
   "split_u0d1:RRTMGX_SW_UP=0,RRTMGX_SW_DOWN=1" --profile > gpurun_out/s2b_sweep.jsonl 2> gpurun_out/s2b_sweep.err
cat gpurun_out/s2b_sweep.jsonl; tail -3 gpurun_out/s2b_sweep.err
