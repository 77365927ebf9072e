#!/bin/bash
# N = 8: bench with the objects dealt out by distance + K1 early exit (parity_check inside the warm-up)
mkdir -p gpurun_out
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 ) > gpurun_out/r03l_bench_n8.json 2> gpurun_out/r03l_bench_n8.err
tail -c 300 gpurun_out/r03l_bench_n8.json; tail -3 gpurun_out/r03l_bench_n8.err
