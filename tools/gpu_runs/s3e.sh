# why the persistent LW kernel fails under debug taps: memcheck on the one test
RRTMGX_LW_GN=4444444444444444 timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest "tests/test_lw_gpu.py::test_lw_mcica_mask_and_cloud_optics" -m gpu -x -q > gpurun_out/s3e_memcheck.log 2>&1
grep -n "Invalid\|Error\|at 0x\|by thread\|Address\|passed\|failed" gpurun_out/s3e_memcheck.log | head -30
