# experiment: first-wave stagger of the SW band kernels (RRTMGX_SW_STAGGER="mode:period_us"): are the resident blocks'
# FP64-heavy upward and DRAM-heavy downward sweeps in lockstep, and does shifting them help?  65 536 columns x L72.
python tools/sweep.py 65536 72 "default:" "m1_300:RRTMGX_SW_STAGGER=1:300" "m1_600:RRTMGX_SW_STAGGER=1:600" \
    "m2_300:RRTMGX_SW_STAGGER=2:300" "m2_600:RRTMGX_SW_STAGGER=2:600" "m1_1200:RRTMGX_SW_STAGGER=1:1200" "default_again:" \
    --profile > gpurun_out/t1h_sweep.jsonl 2> gpurun_out/t1h_sweep.err
python - <<'PY'
import json
for l in open("gpurun_out/t1h_sweep.jsonl"):
    d = json.loads(l)
    print(d["cfg"], "both", round(d["both_ms"], 2), "lw", round(d["lw_ms"], 2), "sw", round(d["sw_ms"], 2),
          {k: v for k, v in list(d["families_ms"].items())[:2]}, d["sum_swdflx"])
PY
tail -3 gpurun_out/t1h_sweep.err
