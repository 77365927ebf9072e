#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_multi_gpu.py -m gpu -x -q > gpurun_out/pytest_mgpu8.log 2>&1; echo "exit $?" >> gpurun_out/pytest_mgpu8.log; tail -3 gpurun_out/pytest_mgpu8.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29668 bench.py --gpus 8 --steps 20 --warmup 3 --workload c5 > gpurun_out/bench_c5_n8_final.json 2> gpurun_out/bench_c5_n8_final.err; tail -c 150 gpurun_out/bench_c5_n8_final.json
