# session 4 of round 2: the SOLAR_RADVAL build on the device + the whole GPU suite in one process, smoke, a short bench line
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/t1a_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/t1a_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/t1a_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/t1a_smoke.log
( time timeout 600 python bench.py --no-cpu --no-e2e --steps 6 --warmup 3 ) > gpurun_out/t1a_bench_short.log 2>&1
tail -25 gpurun_out/t1a_tests.log; tail -2 gpurun_out/t1a_smoke.log; tail -c 600 gpurun_out/t1a_bench_short.log
