# A/B of two SW band-kernel variants against the default build (tools/build_variant.sh; 65 536 columns x L72):
#   keep12 / keep24: the top 12 / 24 layers of the upward sweep's per-cell scratch stored with the default L2 policy instead
#                    of st.cs (the downward sweep reads them back first)
#   red1:            one shared-memory reduction buffer + two barriers per level instead of two buffers + one barrier
#                    (46 instead of 92 KB of shared memory per SM at three blocks: 36 KB more L1)
run() { python tools/sweep.py 65536 72 "$1:" --profile >> gpurun_out/t1d_sweep.jsonl 2>> gpurun_out/t1d_sweep.err; }
: > gpurun_out/t1d_sweep.jsonl
cp geosradiation_gridcomp_b200/librrtmgx.so /tmp/keep.so
run default
for v in keep12 keep24 red1; do
  if [ -f geosradiation_gridcomp_b200/csrc/build_$v/librrtmgx_$v.so ]; then
    cp geosradiation_gridcomp_b200/csrc/build_$v/librrtmgx_$v.so geosradiation_gridcomp_b200/librrtmgx.so
    run $v
  fi
done
cp /tmp/keep.so geosradiation_gridcomp_b200/librrtmgx.so
run default_again
python - <<'PY'
import json
for l in open("gpurun_out/t1d_sweep.jsonl"):
    d = json.loads(l)
    print(d["cfg"], "both", round(d["both_ms"], 2), "lw", round(d["lw_ms"], 2), "sw", round(d["sw_ms"], 2),
          {k: v for k, v in list(d["families_ms"].items())[:3]}, d["sum_swdflx"], d["sum_swuflx"])
PY
