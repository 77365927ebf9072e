#!/bin/bash
mkdir -p gpurun_out
CFG='[{"TGB_K1_KERNEL":1},{"TGB_K1_KERNEL":1,"TGB_K1_MIN_CTAS":5},{"TGB_K1_KERNEL":1,"TGB_K1_MIN_CTAS":6},{"TGB_K1_KERNEL":1,"TGB_K1_MIN_CTAS":3}]'
( time timeout 600 python tools/sweep.py --workload c2 --frames 10 --configs "$CFG" ) > gpurun_out/r02l_sweep_k1.jsonl 2> gpurun_out/r02l_sweep_k1.err; tail -3 gpurun_out/r02l_sweep_k1.err
