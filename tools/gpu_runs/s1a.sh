# round 2, first call: full GPU suite in one process on the re-entry state (reuse fix), smoke, bench
python -m pytest tests -m gpu -x -q > gpurun_out/s1a_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/s1a_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s1a_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/s1a_smoke.log
( time python bench.py ) > gpurun_out/s1a_bench.log 2>&1
nvidia-smi -q | grep -i -E "product name|numa|cpu affinity" | head; lscpu | head -20 > gpurun_out/s1a_lscpu.txt; which gfortran flang ifort nvfortran >> gpurun_out/s1a_lscpu.txt 2>&1
tail -5 gpurun_out/s1a_tests.log; tail -2 gpurun_out/s1a_smoke.log; tail -c 1500 gpurun_out/s1a_bench.log
