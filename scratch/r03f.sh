#!/bin/bash
# exact pool kernel on a tile-sized batch: fair-share rounds (TGB_GI_POOL_ROUNDS), rays per lane
mkdir -p gpurun_out
CFG='[{"TGB_GI_KERNEL":2},{"TGB_GI_POOL_ROUNDS":1},{"TGB_GI_POOL_ROUNDS":2},{"TGB_GI_POOL_ROUNDS":3},{"TGB_GI_POOL_ROUNDS":4},{"TGB_GI_POOL_ROUNDS":2,"TGB_GI_RAYS_PER_LANE":2},{"TGB_GI_POOL_ROUNDS":1,"TGB_GI_RAYS_PER_LANE":2},{"TGB_GI_POOL_ROUNDS":2,"TGB_GI_RAYS_PER_LANE":4},{"TGB_GI_RAYS_PER_LANE":2},{"TGB_GI_RAYS_PER_LANE":4}]'
( time timeout 600 python tools/sweep.py --workload c2 --frames 10 --rows 272 --row0 1088 --configs "$CFG" ) > gpurun_out/r03f_sweep_tile.jsonl 2> gpurun_out/r03f_sweep_tile.err
( time timeout 600 python tools/sweep.py --workload c2 --frames 10 --rows 272 --row0 544 --configs "$CFG" ) > gpurun_out/r03f_sweep_tile2.jsonl 2> gpurun_out/r03f_sweep_tile2.err
tail -2 gpurun_out/r03f_sweep_tile2.err
