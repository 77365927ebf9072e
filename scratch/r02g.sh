#!/bin/bash
# N=2: multi-GPU parity test + the new bench.py path (strong scaling, parity_check, also c5)
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_multi_gpu.py -x -q -m gpu ) > gpurun_out/r02g_pytest_mgpu.log 2>&1; tail -5 gpurun_out/r02g_pytest_mgpu.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 ) > gpurun_out/r02g_bench_n2.json 2> gpurun_out/r02g_bench_n2.err; tail -c 1500 gpurun_out/r02g_bench_n2.json; tail -5 gpurun_out/r02g_bench_n2.err
