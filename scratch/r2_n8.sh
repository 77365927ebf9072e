#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -x -q > gpurun_out/pytest_mgpu8.log 2>&1; echo "exit $?" >> gpurun_out/pytest_mgpu8.log; tail -3 gpurun_out/pytest_mgpu8.log
for n in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2966$n bench.py --gpus $n --steps 20 --warmup 3 --workload c5 > gpurun_out/bench_c5_n$n.json 2> gpurun_out/bench_c5_n$n.err; tail -c 300 gpurun_out/bench_c5_n$n.json; tail -2 gpurun_out/bench_c5_n$n.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2967$n bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/bench_c2_n$n.json 2> gpurun_out/bench_c2_n$n.err; tail -c 300 gpurun_out/bench_c2_n$n.json; tail -2 gpurun_out/bench_c2_n$n.err
done
