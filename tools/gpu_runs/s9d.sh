# LW g-point sums as a shuffle reduce-scatter (same bits, a third of the exchanges): parity, then timing
timeout 600 python -m pytest tests/test_lw_gpu.py tests/test_refexec_pin_gpu.py tests/test_fullsize_gpu.py -q -x > gpurun_out/s9d_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/s9d_tests.log
timeout 300 python tools/sweep.py 65536 72 "rs_shuffle:" "rs_shuffle_again:" --profile > gpurun_out/s9d_sweep.jsonl 2> gpurun_out/s9d_sweep.err
tail -3 gpurun_out/s9d_tests.log; cut -c1-420 gpurun_out/s9d_sweep.jsonl
