# timing only: the same kernels built with --fmad=true (FP64 instruction count down, everything else equal)
python tools/sweep.py 65536 72 "fused_nofma:RRTMGX_SW_SPLIT=0" "split_nofma:RRTMGX_SW_SPLIT=1" --profile > gpurun_out/s2d_sweep.jsonl 2> gpurun_out/s2d_sweep.err
cp geosradiation_gridcomp_b200/librrtmgx.so /tmp/keep.so
cp geosradiation_gridcomp_b200/csrc/build/librrtmgx_fmad.so geosradiation_gridcomp_b200/librrtmgx.so
python tools/sweep.py 65536 72 "fused_fma:RRTMGX_SW_SPLIT=0" "split_fma:RRTMGX_SW_SPLIT=1" --profile >> gpurun_out/s2d_sweep.jsonl 2>> gpurun_out/s2d_sweep.err
cp /tmp/keep.so geosradiation_gridcomp_b200/librrtmgx.so
tail -3 gpurun_out/s2d_sweep.err
