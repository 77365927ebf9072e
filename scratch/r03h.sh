#!/bin/bash
# session 2 check on 2 GPUs: multi-GPU parity test (K1 early exit on the SHARDED path), bench N=2 with parity_check
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_multi_gpu.py -x -q -m gpu ) > gpurun_out/r03h_pytest_mgpu_n2.log 2>&1; tail -3 gpurun_out/r03h_pytest_mgpu_n2.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 ) > gpurun_out/r03h_bench_n2.json 2> gpurun_out/r03h_bench_n2.err
tail -c 400 gpurun_out/r03h_bench_n2.json; tail -3 gpurun_out/r03h_bench_n2.err
