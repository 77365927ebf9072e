#!/bin/bash
# fast GI walk on a tile-sized batch (272 rows = the screen tile of one of 8 ranks) and exact-pass shapes
mkdir -p gpurun_out
CFG='[{"TGB_GI_KERNEL":2},{"TGB_GI_KERNEL":3},{"TGB_GI_KERNEL":3,"TGB_GI_FAST_STEPS":8},{"TGB_GI_KERNEL":3,"TGB_GI_FAST_STEPS":8,"TGB_GI_RAYS_PER_LANE":1,"TGB_GI_POOL_CTAS_PER_SM":8},{"TGB_GI_KERNEL":3,"TGB_GI_FAST_STEPS":8,"TGB_GI_RAYS_PER_LANE":1,"TGB_GI_POOL_CTAS_PER_SM":16},{"TGB_GI_KERNEL":3,"TGB_GI_FAST_STEPS":8,"TGB_GI_RAYS_PER_LANE":1,"TGB_GI_POOL_CTAS_PER_SM":16,"TGB_GI_POOL_SERVICE_SLOTS":8},{"TGB_GI_KERNEL":3,"TGB_GI_FAST_STEPS":8,"TGB_GI_RAYS_PER_LANE":1,"TGB_GI_POOL_CTAS_PER_SM":16,"TGB_GI_POOL_SERVICE_SLOTS":4,"TGB_GI_POOL_DDA_STEPS":32,"TGB_GI_POOL_TREE_REPS":8}]'
( time timeout 600 python tools/sweep.py --workload c2 --frames 10 --rows 272 --row0 1088 --configs "$CFG" ) > gpurun_out/r03d_sweep_tile.jsonl 2> gpurun_out/r03d_sweep_tile.err
( time timeout 600 python tools/sweep.py --workload c2 --frames 10 --configs "$CFG" ) > gpurun_out/r03d_sweep_full.jsonl 2> gpurun_out/r03d_sweep_full.err
tail -2 gpurun_out/r03d_sweep_full.err
