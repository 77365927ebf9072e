#!/bin/bash
# round 2, session 3: the driver's scaling run at N GPUs (argument) with GI kernel 4: bench only
N=${1:-8}
mkdir -p gpurun_out
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 ) > gpurun_out/r04i_bench_n$N.json 2> gpurun_out/r04i_bench_n$N.err
tail -c 400 gpurun_out/r04i_bench_n$N.json; tail -3 gpurun_out/r04i_bench_n$N.err
