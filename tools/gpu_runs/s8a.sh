# the CUDA path against the golden vectors made from the reference's own source (tests/test_refexec_pin_gpu.py)
python -m pytest tests/test_refexec_pin_gpu.py -q -s > gpurun_out/s8a_refexec_pin.log 2>&1; echo "tests rc=$?" >> gpurun_out/s8a_refexec_pin.log
tail -30 gpurun_out/s8a_refexec_pin.log
