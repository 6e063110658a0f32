# LW band kernels: Planck fractions kept per cell, pipelined look-ups in the upward sweep, shuffle sums: parity + timing
python -m pytest tests/test_lw_gpu.py tests/test_fullsize_gpu.py tests/test_aa_regress_gpu.py tests/test_zz_options_gpu.py -m gpu -x -q > gpurun_out/s3a_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/s3a_tests.log
tail -4 gpurun_out/s3a_tests.log
python tools/sweep.py 65536 72 "default:RRTMGX_SW_SPLIT=0" "cb8_b35:RRTMGX_LW_GN=0123223232222111" "cb16_pow2:RRTMGX_LW_GN=0113113132111111" --profile --only=lw > gpurun_out/s3a_sweep.jsonl 2> gpurun_out/s3a_sweep.err
tail -3 gpurun_out/s3a_sweep.err
