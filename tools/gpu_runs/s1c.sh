# full GPU suite in one process after the refused-call early exit
python -m pytest tests -m gpu -x -q > gpurun_out/s1c_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/s1c_tests.log
tail -5 gpurun_out/s1c_tests.log
