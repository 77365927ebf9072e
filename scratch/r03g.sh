#!/bin/bash
# K1: untouched tiles leave after one barrier -- full scene and the shards of ranks 0 / 6 of 8 on one GPU
mkdir -p gpurun_out
CFG='[{"TGB_K1_EARLY_EXIT":0},{"TGB_K1_EARLY_EXIT":1}]'
for sh in 0,1 0,8 6,8 7,8 3,4; do
  ( timeout 300 python tools/sweep.py --workload c2 --frames 12 --what k1 --shard $sh --configs "$CFG" ) > gpurun_out/r03g_sweep_k1_${sh/,/of}.jsonl 2> gpurun_out/r03g_sweep_k1_${sh/,/of}.err
done
( timeout 300 python tools/sweep.py --workload c2far --frames 12 --what k1 --configs "$CFG" ) > gpurun_out/r03g_sweep_k1_far.jsonl 2> gpurun_out/r03g_sweep_k1_far.err
( time timeout 900 python -m pytest tests/test_visibility_gpu.py -m gpu -x -q ) > gpurun_out/r03g_pytest_vis.log 2>&1; tail -3 gpurun_out/r03g_pytest_vis.log
