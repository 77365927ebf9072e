#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -x -q -s > gpurun_out/pytest_mgpu8.log 2>&1; echo "exit $?" >> gpurun_out/pytest_mgpu8.log; tail -4 gpurun_out/pytest_mgpu8.log
run() { n=$1; w=$2; m=$3; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2966$n bench.py --gpus $n --steps 20 --warmup 3 --workload $w --merge $m > gpurun_out/bench_${w}_n${n}_$m.json 2> gpurun_out/bench_${w}_n${n}_$m.err; tail -c 150 gpurun_out/bench_${w}_n${n}_$m.json; grep -v "^\*\*\*\|OMP_NUM" gpurun_out/bench_${w}_n${n}_$m.err | tail -3; }
run 8 c5 peer
run 8 c5 nccl
run 8 c2 peer
run 4 c2 peer
