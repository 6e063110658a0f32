# SW register budgets re-measured after the tiling (v0 tuned, v1 48, v2 72 (64 where tuned >= 72), v3 56 <-> 64)
for v in 0 1 2 3; do
  RRTMGX_SW_GN=$v python tools/profile_step.py 65536 72 2 > gpurun_out/r3e_prof_sw$v.json 2> gpurun_out/r3e_prof_sw$v.err
done
