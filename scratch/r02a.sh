#!/bin/bash
# round 2, first GPU call: full GPU suite on the new tree, GI kernel sweep (pool vs flat vs stack), default bench (timed)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02a_smi.txt
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02a_pytest_gpu.log 2>&1; echo "exit $?" >> gpurun_out/r02a_pytest_gpu.log; tail -5 gpurun_out/r02a_pytest_gpu.log
( time timeout 600 python tools/sweep.py --workload c2 --what gi --frames 10 ) > gpurun_out/r02a_sweep_gi.jsonl 2> gpurun_out/r02a_sweep_gi.err; tail -3 gpurun_out/r02a_sweep_gi.err
( time timeout 300 python tools/sweep.py --workload c2 --what gi --frames 10 --rows 270 --configs '[{"TGB_GI_KERNEL":1},{"TGB_GI_KERNEL":2},{"TGB_GI_KERNEL":2,"TGB_GI_POOL_MIN_RAYS_PER_SLOT":1},{"TGB_GI_KERNEL":2,"TGB_GI_POOL_MIN_RAYS_PER_SLOT":2},{"TGB_GI_KERNEL":2,"TGB_GI_POOL_MIN_RAYS_PER_SLOT":8},{"TGB_GI_KERNEL":2,"TGB_GI_POOL_MIN_RAYS_PER_SLOT":16},{"TGB_GI_KERNEL":2,"TGB_GI_RAYS_PER_LANE":2,"TGB_GI_POOL_MIN_RAYS_PER_SLOT":2},{"TGB_GI_KERNEL":2,"TGB_GI_RAYS_PER_LANE":2,"TGB_GI_POOL_MIN_RAYS_PER_SLOT":8}]' ) > gpurun_out/r02a_sweep_gi_tile.jsonl 2> gpurun_out/r02a_sweep_gi_tile.err
( time timeout 900 python bench.py ) > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err; tail -c 600 gpurun_out/r02a_bench.json; tail -5 gpurun_out/r02a_bench.err
