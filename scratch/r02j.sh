#!/bin/bash
# N=8: multi-GPU parity test + bench (c2 strong scaling + c5 under `also`)
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_multi_gpu.py -x -q -m gpu ) > gpurun_out/r02j_pytest_mgpu8.log 2>&1; tail -6 gpurun_out/r02j_pytest_mgpu8.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 3 ) > gpurun_out/r02j_bench_n8.json 2> gpurun_out/r02j_bench_n8.err; tail -c 400 gpurun_out/r02j_bench_n8.json; tail -5 gpurun_out/r02j_bench_n8.err
