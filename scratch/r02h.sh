#!/bin/bash
# N=2: multi-GPU parity test + bench after the merge rewrite (tile flags, K1 epilogue materials, device-side barrier)
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_multi_gpu.py -x -q -m gpu ) > gpurun_out/r02h_pytest_mgpu.log 2>&1; tail -15 gpurun_out/r02h_pytest_mgpu.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 ) > gpurun_out/r02h_bench_n2.json 2> gpurun_out/r02h_bench_n2.err; tail -c 600 gpurun_out/r02h_bench_n2.json; tail -5 gpurun_out/r02h_bench_n2.err
