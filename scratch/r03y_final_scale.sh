#!/bin/bash
# the driver's scaling run, reproduced: bench.py at N GPUs (argument), both arms
N=$1
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  ( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r03y_bench_n1.json 2> gpurun_out/r03y_bench_n1.err
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gi_trace_pool --launch-skip 3 -c 1 -o gpurun_out/r03y_k3b -f python tools/sweep.py --workload c2 --frames 3 --configs '[{}]' > gpurun_out/r03y_k3b.log 2>&1
else
  ( time timeout 900 python -m pytest tests/test_multi_gpu.py -x -q -m gpu ) > gpurun_out/r03y_pytest_mgpu_n$N.log 2>&1; tail -3 gpurun_out/r03y_pytest_mgpu_n$N.log
  ( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 ) > gpurun_out/r03y_bench_n$N.json 2> gpurun_out/r03y_bench_n$N.err
  ( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 --merge nccl --also c5 ) > gpurun_out/r03y_bench_n${N}_nccl.json 2> gpurun_out/r03y_bench_n${N}_nccl.err
fi
tail -c 300 gpurun_out/r03y_bench_n$N.json; tail -3 gpurun_out/r03y_bench_n$N.err
