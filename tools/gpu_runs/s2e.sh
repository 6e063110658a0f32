# SW band kernels with row-pair tables, fma / reciprocal quotients, own exp: parity tests, then timing fused vs split
python -m pytest tests/test_sw_gpu.py tests/test_fullsize_gpu.py tests/test_glue_gpu.py tests/test_zz_options_gpu.py -m gpu -x -q > gpurun_out/s2e_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/s2e_tests.log
tail -4 gpurun_out/s2e_tests.log
python tools/sweep.py 65536 72 "fused:RRTMGX_SW_SPLIT=0" "split_u0:RRTMGX_SW_SPLIT=1,RRTMGX_SW_UP=0" "split_u1:RRTMGX_SW_SPLIT=1,RRTMGX_SW_UP=1" --profile > gpurun_out/s2e_sweep.jsonl 2> gpurun_out/s2e_sweep.err
tail -3 gpurun_out/s2e_sweep.err
