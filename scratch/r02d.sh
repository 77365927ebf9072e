#!/bin/bash
mkdir -p gpurun_out
run() { # tag, kernel regex, config json
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 --launch-skip 3 -c 1 -o gpurun_out/r02d_$1 -f python tools/sweep.py --workload c2 --frames 3 --configs "$3" > gpurun_out/r02d_$1.log 2>&1
  tail -2 gpurun_out/r02d_$1.log
}
run k1pool_k2 k_visibility_pool '[{"TGB_K1_KERNEL":2,"TGB_K1_PIXELS_PER_LANE":2}]'
run k1pool_k1 k_visibility_pool '[{"TGB_K1_KERNEL":2,"TGB_K1_PIXELS_PER_LANE":1}]'
ls -la gpurun_out/r02d*.ncu-rep
