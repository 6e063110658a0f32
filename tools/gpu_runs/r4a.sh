# per-GPU slabs of BASELINE.json configs 4 and 5 on one B200 (weak scaling is 8.0x, profiles/r2_q_bench_8gpu.json):
#   C720 x L72  = 3 110 400 columns / 8 GPUs = 388 800 columns per GPU
#   C360 x L181 =   777 600 columns / 8 GPUs =  97 200 columns per GPU
( time python bench.py --ncol 388800 --nlay 72 --steps 3 --warmup 3 --no-cpu ) > gpurun_out/r4a_c720_slab.log 2>&1
( time python bench.py --ncol 97200 --nlay 181 --steps 3 --warmup 3 --no-cpu ) > gpurun_out/r4a_c360l181_slab.log 2>&1
tail -c 600 gpurun_out/r4a_c720_slab.log; tail -c 600 gpurun_out/r4a_c360l181_slab.log
