#!/bin/bash
# pool kernel: longer phases for the last few rays of a warp (TGB_GI_POOL_TAIL_BOOST) on tile-sized batches and the full frame
mkdir -p gpurun_out
CFG='[{},{"TGB_GI_POOL_TAIL_BOOST":2},{"TGB_GI_POOL_TAIL_BOOST":4},{"TGB_GI_POOL_TAIL_BOOST":8},{"TGB_GI_POOL_TAIL_BOOST":16}]'
( timeout 600 python tools/sweep.py --workload c2 --frames 12 --rows 272 --row0 1088 --configs "$CFG" ) > gpurun_out/r03q_sweep_tile.jsonl 2> gpurun_out/r03q_sweep_tile.err
( timeout 600 python tools/sweep.py --workload c2 --frames 12 --rows 272 --row0 544 --configs "$CFG" ) > gpurun_out/r03q_sweep_tile2.jsonl 2> gpurun_out/r03q_sweep_tile2.err
( timeout 600 python tools/sweep.py --workload c2 --frames 12 --configs "$CFG" ) > gpurun_out/r03q_sweep_full.jsonl 2> gpurun_out/r03q_sweep_full.err
