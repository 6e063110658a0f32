# RRTMGX_LIT_ONLY (daytime packing in the Solar glue): new tests, then the glue tests that share the kernels
python -m pytest tests/test_litonly_gpu.py tests/test_glue_gpu.py -q -x > gpurun_out/s8b_litonly.log 2>&1; echo "tests rc=$?" >> gpurun_out/s8b_litonly.log
tail -30 gpurun_out/s8b_litonly.log
