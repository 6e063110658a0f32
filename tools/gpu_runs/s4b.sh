# two GPUs: the multi-rank hardware test (NCCL gather of slab fluxes == single-rank recomputation), host bandwidth with two ranks
python -m pytest tests/test_multigpu_gpu.py -m gpu -x -q > gpurun_out/s4b_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/s4b_tests.log
tail -4 gpurun_out/s4b_tests.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/host_bw_probe.py > gpurun_out/s4b_hostbw_2.json 2> gpurun_out/s4b_hostbw.err
cat gpurun_out/s4b_hostbw_2.json
