# experiment: ddiv without the second Newton step of the reciprocal (6 instead of 8 fp64 instructions per quotient).
# Timing of the default build, then the variant build swapped in: timing, the IEEE division pin, the reference-source pin.
python tools/sweep.py 65536 72 "default:" --profile > gpurun_out/s8c_sweep.jsonl 2> gpurun_out/s8c_sweep.err
cp geosradiation_gridcomp_b200/librrtmgx.so /tmp/keep.so
cp geosradiation_gridcomp_b200/csrc/build_short/librrtmgx_short.so geosradiation_gridcomp_b200/librrtmgx.so
python tools/sweep.py 65536 72 "ddiv_short:" --profile >> gpurun_out/s8c_sweep.jsonl 2>> gpurun_out/s8c_sweep.err
python -m pytest tests/test_lw_gpu.py -q -k "divi" > gpurun_out/s8c_tests.log 2>&1
python -m pytest tests/test_refexec_pin_gpu.py tests/test_sw_gpu.py -q >> gpurun_out/s8c_tests.log 2>&1
cp /tmp/keep.so geosradiation_gridcomp_b200/librrtmgx.so
python tools/sweep.py 65536 72 "default_again:" --profile >> gpurun_out/s8c_sweep.jsonl 2>> gpurun_out/s8c_sweep.err
cat gpurun_out/s8c_sweep.jsonl | cut -c1-600; tail -5 gpurun_out/s8c_tests.log
