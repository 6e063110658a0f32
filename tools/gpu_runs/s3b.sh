# LW persistent kernel with the transmittance table in shared memory (bulk copy) + warp-shuffle sums: parity, timing
RRTMGX_LW_GN=4444444444444444 python -m pytest tests/test_lw_gpu.py tests/test_fullsize_gpu.py -m gpu -x -q > gpurun_out/s3b_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/s3b_tests.log
tail -4 gpurun_out/s3b_tests.log
python tools/sweep.py 65536 72 "default:RRTMGX_SW_SPLIT=0" "persist24_all:RRTMGX_LW_GN=4444444444444444" "persist16_all:RRTMGX_LW_GN=5555555555555555" "persist24_big:RRTMGX_LW_GN=4444424232222111" "persist16_big:RRTMGX_LW_GN=5555525232222111" --profile --only=lw > gpurun_out/s3b_sweep.jsonl 2> gpurun_out/s3b_sweep.err
tail -3 gpurun_out/s3b_sweep.err
python tools/profile_step.py 65536 72 1 > gpurun_out/s3b_prof_default.json 2>/dev/null
RRTMGX_LW_GN=4444444444444444 python tools/profile_step.py 65536 72 1 > gpurun_out/s3b_prof_p24.json 2>/dev/null
RRTMGX_LW_GN=5555555555555555 python tools/profile_step.py 65536 72 1 > gpurun_out/s3b_prof_p16.json 2>/dev/null
