# LW 16-byte table gathers; SW downward-sweep L2 prefetch look-ahead 4 (default build) against 0 and 8
set -x
P=geosradiation_gridcomp_b200
python -m pytest tests -m gpu -x -q > gpurun_out/r2e_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2e_tests.log
python tools/profile_step.py 32768 72 2 > gpurun_out/r2e_prof_pf4.json 2> gpurun_out/r2e_prof.err
cp $P/librrtmgx.so /tmp/keep.so
for v in 0 8; do
  cp $P/librrtmgx_pf$v.so $P/librrtmgx.so
  python tools/profile_step.py 32768 72 2 > gpurun_out/r2e_prof_pf$v.json 2>> gpurun_out/r2e_prof.err
done
cp /tmp/keep.so $P/librrtmgx.so
python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r2e_bench.log 2>&1
