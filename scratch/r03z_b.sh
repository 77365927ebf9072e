#!/bin/bash
# the ncu captures the first script missed (ncu matches the base name now)
mkdir -p gpurun_out
run() { # tag, kernel name, workload, skip
  timeout 600 ncu --set full --clock-control none --import-source on -k $2 --launch-skip $4 -c 1 -o gpurun_out/r03z_$1 -f python tools/sweep.py --workload $3 --frames 3 --configs '[{}]' > gpurun_out/r03z_$1.log 2>&1
}
run k1 k_visibility c2 3
run k1_far k_visibility c2far 3
run k2 k_svo_fill_leaves c2 0
ls -la gpurun_out/r03z_*.ncu-rep
