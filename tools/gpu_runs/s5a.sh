# eight GPUs: BASELINE.json configs 4 (C720 x L72) and 5 (C360 x L181) cut into one slab per GPU, with the NCCL verification;
# host -> device bandwidth of the box with eight ranks pulling at once
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
( time run 29521 bench.py --gpus 8 --config 4 --steps 3 --warmup 3 --no-e2e --no-cpu ) > gpurun_out/s5a_config4.log 2>&1
( time run 29522 bench.py --gpus 8 --config 5 --steps 3 --warmup 3 --no-e2e --no-cpu ) > gpurun_out/s5a_config5.log 2>&1
run 29523 tools/host_bw_probe.py > gpurun_out/s5a_hostbw_8.json 2> gpurun_out/s5a_hostbw.err
free -g > gpurun_out/s5a_free.txt; nproc >> gpurun_out/s5a_free.txt
grep -h '^{' gpurun_out/s5a_config4.log gpurun_out/s5a_config5.log | cut -c1-600; cat gpurun_out/s5a_hostbw_8.json
