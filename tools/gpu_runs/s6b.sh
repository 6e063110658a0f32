# cap the resident SW band blocks per SM (unused dynamic shared memory) so LW blocks co-reside: does the step overlap?
python tools/sweep.py 65536 72 "base:RRTMGX_SW_SPLIT=0" "smem100k:RRTMGX_SW_RESERVE_SMEM=102400" "smem70k:RRTMGX_SW_RESERVE_SMEM=71680" "smem60k:RRTMGX_SW_RESERVE_SMEM=61440" "smem45k:RRTMGX_SW_RESERVE_SMEM=46080" > gpurun_out/s6b_sweep.jsonl 2> gpurun_out/s6b_sweep.err
tail -2 gpurun_out/s6b_sweep.err
