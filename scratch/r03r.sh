#!/bin/bash
# the GI kernel on a tile-sized batch (272 rows = what a rank of 8 shades): ncu capture for the N = 8 analysis
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k k_gi_trace_pool --launch-skip 3 -c 1 -o gpurun_out/r03r_k3b_tile -f python tools/sweep.py --workload c2 --frames 3 --rows 272 --row0 544 --configs '[{}]' > gpurun_out/r03r_k3b_tile.log 2>&1
ls -la gpurun_out/r03r_*
