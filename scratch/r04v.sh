#!/bin/bash
# round 2, session 3: queue chunk per warp on tile-sized batches (and on the whole frame)
mkdir -p gpurun_out
CFG='[{"TGB_GI_KERNEL":2},{},{"TGB_GI_FAST_CHUNK":64},{"TGB_GI_FAST_CHUNK":32},{"TGB_GI_FAST_CHUNK":16},{"TGB_GI_FAST_CHUNK":32,"TGB_GI_FAST_SERVICE_LANES":4},{"TGB_GI_FAST_CHUNK":16,"TGB_GI_FAST_SERVICE_LANES":4},{},{"TGB_GI_FAST_CHUNK":64}]'
for R0 in 944 400; do
( time timeout 300 python tools/sweep.py --workload c2 --frames 12 --rows 272 --row0 $R0 --configs "$CFG" ) > gpurun_out/r04v_sweep_tile_$R0.jsonl 2> gpurun_out/r04v_sweep_tile_$R0.err
done
( time timeout 300 python tools/sweep.py --workload c2 --frames 12 --configs '[{"TGB_GI_KERNEL":2},{},{"TGB_GI_FAST_CHUNK":32},{"TGB_GI_FAST_CHUNK":128},{}]' ) > gpurun_out/r04v_sweep_full.jsonl 2> gpurun_out/r04v_sweep_full.err
