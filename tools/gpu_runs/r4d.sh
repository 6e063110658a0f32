# staging depth of the host-array pipeline (2 = the old double buffering, 3 = new default, 4), then the GPU suite on the new default
PROBE_CHUNKS=8192,16384 PROBE_STAGES=2,3,4 PROBE_MODES=1 PROBE_STEPS=6 python tools/e2e_probe.py > gpurun_out/r4d_e2e_stages.jsonl 2> gpurun_out/r4d_e2e_stages.err
python -m pytest tests -m gpu -x -q > gpurun_out/r4d_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r4d_tests.log
cat gpurun_out/r4d_e2e_stages.jsonl; tail -3 gpurun_out/r4d_e2e_stages.err; tail -3 gpurun_out/r4d_tests.log
