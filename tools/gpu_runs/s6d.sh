# LW: whole-warp blocks for the bands whose tuned blocks are 24 / 28 threads (bands 4, 7, 9, 10); real*4 device arrays test
python -m pytest tests/test_hostpipe_gpu.py -m gpu -x -q > gpurun_out/s6d_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/s6d_tests.log; tail -3 gpurun_out/s6d_tests.log
python tools/sweep.py 65536 72 "default:RRTMGX_LW_GN=0133323232222111" "b7b9_cb16:RRTMGX_LW_GN=0133321212222111" "b4b10_cb32:RRTMGX_LW_GN=0130323230222111" "all4:RRTMGX_LW_GN=0130321210222111" "b4_cb8_b10_cb16:RRTMGX_LW_GN=0132323231222111" --only=lw > gpurun_out/s6d_sweep.jsonl 2> gpurun_out/s6d_sweep.err
python tools/profile_step.py 65536 72 1 > gpurun_out/s6d_prof_default.json 2>/dev/null
RRTMGX_LW_GN=0130321210222111 python tools/profile_step.py 65536 72 1 > gpurun_out/s6d_prof_all4.json 2>/dev/null
RRTMGX_LW_GN=0132323231222111 python tools/profile_step.py 65536 72 1 > gpurun_out/s6d_prof_mid.json 2>/dev/null
tail -2 gpurun_out/s6d_sweep.err
