# SW upward kernel with two g-points per thread (96 / 128 registers) against one g-point per thread
python tools/sweep.py 65536 72 "split_u1d0:RRTMGX_SW_UP=1,RRTMGX_SW_DOWN=0" "split_u3d0:RRTMGX_SW_UP=3,RRTMGX_SW_DOWN=0" "split_u4d0:RRTMGX_SW_UP=4,RRTMGX_SW_DOWN=0" --profile --only=sw > gpurun_out/s2c_sweep.jsonl 2> gpurun_out/s2c_sweep.err
cat gpurun_out/s2c_sweep.jsonl; tail -3 gpurun_out/s2c_sweep.err
