# host staging chunk against the three end-to-end arms of bench.py (fp64 arrays, real*4 arrays, native real*4 glue):
# with half the bytes the kernels, not the link, are the longer stage, and an 8192-column chunk does not fill the SMs
for c in 8192 16384 32768 65536; do
  RRTMGX_HOST_CHUNK=$c timeout 300 python bench.py --no-cpu --steps 4 --warmup 3 > gpurun_out/t1b_bench_chunk$c.log 2>&1
  python - <<PY
import json
for line in open("gpurun_out/t1b_bench_chunk$c.log"):
    if line.startswith("{"):
        d = json.loads(line); e = d["e2e"]
        print(json.dumps({"host_chunk": $c, "device_resident": round(d["value"]), "e2e_fp64_arrays": round(e["value"]),
                          "e2e_real4_arrays": round(e["real4_host_arrays"]["value"]),
                          "e2e_native_real4_glue": e["native_real4_glue"].get("value") and round(e["native_real4_glue"]["value"]),
                          "sm_mhz": d["clocks"]["sm_mhz"]}))
PY
done | tee gpurun_out/t1b_e2e_chunk_sweep.jsonl
