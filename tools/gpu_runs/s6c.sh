# SW fused kernel: where the downward sweep's scratch prefetch goes (L1 one level ahead = default, none, L2 one / two levels ahead)
python tools/sweep.py 65536 72 "pf_l1:RRTMGX_SW_PF=0" "pf_none:RRTMGX_SW_PF=1" "pf_l2:RRTMGX_SW_PF=2" "pf_l2x2:RRTMGX_SW_PF=3" "pf_l2x3:RRTMGX_SW_PF=4" --profile > gpurun_out/s6c_sweep.jsonl 2> gpurun_out/s6c_sweep.err
tail -2 gpurun_out/s6c_sweep.err
