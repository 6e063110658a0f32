# LW (g-points per thread, registers) re-measured at the tuned CB: v0 tuned, v1 other GN r64, v2 r80, v3 r56
for v in 0 1 2 3; do
  RRTMGX_LW_GN=$v python tools/profile_step.py 65536 72 2 > gpurun_out/r3c_prof_lw$v.json 2> gpurun_out/r3c_prof_lw$v.err
done
