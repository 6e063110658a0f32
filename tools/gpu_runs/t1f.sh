# final state of session 4: whole GPU suite in one process, smoke, and the device cost of the SOLAR_RADVAL build
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t1f_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/t1f_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/t1f_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/t1f_smoke.log
timeout 300 python tools/radval_cost.py 65536 > gpurun_out/t1f_radval_cost.jsonl 2> gpurun_out/t1f_radval_cost.err
tail -4 gpurun_out/t1f_tests.log; tail -2 gpurun_out/t1f_smoke.log; cat gpurun_out/t1f_radval_cost.jsonl; tail -3 gpurun_out/t1f_radval_cost.err
