# full GPU suite in one process after the NaN-trap fix; FP64 issue roof probe
python -m pytest tests -m gpu -x -q > gpurun_out/s1b_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/s1b_tests.log
python tools/fp64_peak.py gpurun_out/fp64_peak.json > gpurun_out/s1b_fp64.log 2>&1
tail -5 gpurun_out/s1b_tests.log; cat gpurun_out/s1b_fp64.log
