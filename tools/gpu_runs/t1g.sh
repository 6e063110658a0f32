# experiment: preferred shared-memory carve-out of the band kernels (RRTMGX_CARVEOUT, percent of the SM's 256 KB; the
# rest is L1).  One process per setting (the attribute is set at a kernel's first launch); 65 536 columns x L72.
: > gpurun_out/t1g_sweep.jsonl
for c in "" 0 25 40 50 100; do
  if [ -z "$c" ]; then name=driver; unset RRTMGX_CARVEOUT; else name=carve$c; export RRTMGX_CARVEOUT=$c; fi
  python tools/sweep.py 65536 72 "$name:" --profile >> gpurun_out/t1g_sweep.jsonl 2>> gpurun_out/t1g_sweep.err
done
unset RRTMGX_CARVEOUT
python tools/sweep.py 65536 72 "driver_again:" --profile >> gpurun_out/t1g_sweep.jsonl 2>> gpurun_out/t1g_sweep.err
python - <<'PY'
import json
for l in open("gpurun_out/t1g_sweep.jsonl"):
    d = json.loads(l)
    print(d["cfg"], "both", round(d["both_ms"], 2), "lw", round(d["lw_ms"], 2), "sw", round(d["sw_ms"], 2),
          {k: v for k, v in list(d["families_ms"].items())[:3]}, d["sum_swdflx"])
PY
