# timing only: N extra independent integer instructions per SW cell (is the band kernel bound by its issue rate?)
python tools/sweep.py 65536 72 "pad0:RRTMGX_SW_SPLIT=0" --profile --only=sw > gpurun_out/s6a_sweep.jsonl 2> gpurun_out/s6a_sweep.err
cp geosradiation_gridcomp_b200/librrtmgx.so /tmp/keep.so
for n in 80 160; do
  cp geosradiation_gridcomp_b200/csrc/build/librrtmgx_pad$n.so geosradiation_gridcomp_b200/librrtmgx.so
  python tools/sweep.py 65536 72 "pad$n:RRTMGX_SW_SPLIT=0" --profile --only=sw >> gpurun_out/s6a_sweep.jsonl 2>> gpurun_out/s6a_sweep.err
done
cp /tmp/keep.so geosradiation_gridcomp_b200/librrtmgx.so
tail -2 gpurun_out/s6a_sweep.err
