set -x
P=geosradiation_gridcomp_b200
python -m pytest tests -m gpu -x -q > gpurun_out/r2a_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2a_tests.log
for v in 0 1 2 3; do
  RRTMGX_LW_GN=$v RRTMGX_SW_GN=$v python tools/profile_step.py 32768 72 2 > gpurun_out/r2a_prof_v$v.json 2> gpurun_out/r2a_prof_v$v.err
done
for v in 2 3; do
  RRTMGX_LW_GN=$v RRTMGX_SW_GN=$v python -m pytest tests/test_lw_gpu.py tests/test_sw_gpu.py -m gpu -q > gpurun_out/r2a_tests_v$v.log 2>&1
done
cp $P/librrtmgx.so /tmp/keep.so; cp $P/librrtmgx_fma.so $P/librrtmgx.so
python -m pytest tests/test_lw_gpu.py tests/test_sw_gpu.py -m gpu -q > gpurun_out/r2a_tests_fma.log 2>&1
for v in 0 2; do
  RRTMGX_LW_GN=$v RRTMGX_SW_GN=$v python tools/profile_step.py 32768 72 2 > gpurun_out/r2a_prof_fma_v$v.json 2> gpurun_out/r2a_prof_fma_v$v.err
done
cp /tmp/keep.so $P/librrtmgx.so
