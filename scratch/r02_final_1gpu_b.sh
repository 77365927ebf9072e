#!/bin/bash
mkdir -p gpurun_out
run() { # tag, kernel regex, workload, config json, skip
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 --launch-skip $5 -c 1 -o gpurun_out/r02z_$1 -f python tools/sweep.py --workload $3 --frames 3 --configs "$4" > gpurun_out/r02z_$1.log 2>&1
}
run k1 '^k_visibility$' c2 '[{}]' 3
run k1_far '^k_visibility$' c2far '[{}]' 3
run k2 k_svo_fill_leaves c2 '[{}]' 0
ls -la gpurun_out/r02z_*.ncu-rep
