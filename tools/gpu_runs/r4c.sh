# host staging chunk sweep of the end-to-end path (raw H2D of the box: 55.6 GB/s, gpurun_out/r4b_pcie.json)
python tools/e2e_probe.py > gpurun_out/r4c_e2e_probe.jsonl 2> gpurun_out/r4c_e2e_probe.err
cat gpurun_out/r4c_e2e_probe.jsonl; tail -3 gpurun_out/r4c_e2e_probe.err
