#!/bin/bash
mkdir -p gpurun_out
CFG='[{"TGB_GI_KERNEL":1},{"TGB_GI_POOL_GRID16":0},{"TGB_GI_POOL_GRID16":1},{"TGB_GI_POOL_GRID16":1,"TGB_GI_RAYS_PER_LANE":2},{"TGB_GI_POOL_GRID16":1,"TGB_GI_RAYS_PER_LANE":4},{"TGB_GI_POOL_GRID16":1,"TGB_GI_RAYS_PER_LANE":4,"TGB_GI_POOL_SERVICE_SLOTS":80}]'
( time timeout 600 python tools/sweep.py --workload c2 --frames 10 --configs "$CFG" ) > gpurun_out/r02p_sweep_gi.jsonl 2> gpurun_out/r02p_sweep_gi.err; tail -3 gpurun_out/r02p_sweep_gi.err
