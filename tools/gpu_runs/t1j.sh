# experiment: the device-resident step with LW and SW on two streams at once (default) against SW after LW, alternated
# A B A B A B in one call (box drift), full C180 grid, 6 steps each
for i in 1 2 3; do
  for m in concurrent serial; do
    timeout 300 python bench.py --no-cpu --no-e2e --verify-cols 0 --steps 6 --warmup 3 --paths $m > gpurun_out/t1j_$m$i.log 2>&1
    python - <<PY
import json
for line in open("gpurun_out/t1j_$m$i.log"):
    if line.startswith("{"):
        d = json.loads(line)
        print(json.dumps({"paths": "$m", "run": $i, "columns_per_s": round(d["value"]), "ms_per_step": round(d["ms_per_step"], 2), "sm_mhz": d["clocks"]["sm_mhz"]}))
PY
  done
done | tee gpurun_out/t1j_paths.jsonl
