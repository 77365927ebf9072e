#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -4
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -x -q > gpurun_out/pytest_mgpu.log 2>&1; echo "exit $?" >> gpurun_out/pytest_mgpu.log; tail -4 gpurun_out/pytest_mgpu.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_c2_n2.json 2> gpurun_out/bench_c2_n2.err; tail -c 400 gpurun_out/bench_c2_n2.json; tail -3 gpurun_out/bench_c2_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29656 bench.py --gpus 2 --steps 20 --warmup 3 --workload c5 --no-cpu-baseline > gpurun_out/bench_c5_n2.json 2> gpurun_out/bench_c5_n2.err; tail -c 400 gpurun_out/bench_c5_n2.json; tail -3 gpurun_out/bench_c5_n2.err
