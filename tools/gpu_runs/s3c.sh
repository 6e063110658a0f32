# ablation timing (results are wrong by construction): what the SW scratch traffic and the LW table look-ups cost
python tools/sweep.py 65536 72 "fused:RRTMGX_SW_SPLIT=0" "no_stores:RRTMGX_SW_SPLIT=0,RRTMGX_ABLATE=1" "no_loads:RRTMGX_SW_SPLIT=0,RRTMGX_ABLATE=2" "neither:RRTMGX_SW_SPLIT=0,RRTMGX_ABLATE=3" --profile --only=sw > gpurun_out/s3c_sweep.jsonl 2> gpurun_out/s3c_sweep.err
python tools/sweep.py 65536 72 "lw:RRTMGX_SW_SPLIT=0" "lw_lookup0:RRTMGX_ABLATE=4" --profile --only=lw >> gpurun_out/s3c_sweep.jsonl 2>> gpurun_out/s3c_sweep.err
tail -3 gpurun_out/s3c_sweep.err
