# the GPU suite on the committed state (oracle gained test hooks since r3h), smoke, and the raw PCIe rates of the box
python -m pytest tests -m gpu -x -q > gpurun_out/r4b_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r4b_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r4b_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r4b_smoke.log
python tools/pcie_probe.py > gpurun_out/r4b_pcie.json 2> gpurun_out/r4b_pcie.err
tail -3 gpurun_out/r4b_tests.log; tail -2 gpurun_out/r4b_smoke.log; cat gpurun_out/r4b_pcie.json
