#!/bin/bash
mkdir -p gpurun_out
export TGB200_VERBOSE=1
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -x -q > gpurun_out/pytest_mgpu.log 2>&1; echo "exit $?" >> gpurun_out/pytest_mgpu.log; tail -25 gpurun_out/pytest_mgpu.log
for m in peer nccl; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 2 --steps 20 --warmup 3 --merge $m > gpurun_out/bench_c2_n2_$m.json 2> gpurun_out/bench_c2_n2_$m.err; tail -c 300 gpurun_out/bench_c2_n2_$m.json; grep -v "^\*\*\*\|OMP_NUM" gpurun_out/bench_c2_n2_$m.err | tail -5
done
