#!/bin/bash
# K1 pool kernel: parity (GPU suite with the pool kernel as default) + sweep on c2 and c2far
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02c_pytest_gpu.log 2>&1; echo "exit $?" >> gpurun_out/r02c_pytest_gpu.log; tail -5 gpurun_out/r02c_pytest_gpu.log
( time timeout 600 python tools/sweep.py --workload c2 --what k1 --frames 10 ) > gpurun_out/r02c_sweep_k1.jsonl 2> gpurun_out/r02c_sweep_k1.err; tail -3 gpurun_out/r02c_sweep_k1.err
( time timeout 600 python tools/sweep.py --workload c2far --what k1 --frames 10 ) > gpurun_out/r02c_sweep_k1_far.jsonl 2> gpurun_out/r02c_sweep_k1_far.err; tail -3 gpurun_out/r02c_sweep_k1_far.err
