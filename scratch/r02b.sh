#!/bin/bash
# ncu --set full of the pool kernel in two configurations (one launch each)
mkdir -p gpurun_out
run() { # tag, kernel regex, config json
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 --launch-skip 3 -c 1 -o gpurun_out/r02b_$1 -f python tools/sweep.py --workload c2 --frames 3 --configs "$3" > gpurun_out/r02b_$1.log 2>&1
  tail -2 gpurun_out/r02b_$1.log
}
run pool_k4 k_gi_trace_pool '[{"TGB_GI_KERNEL":2,"TGB_GI_RAYS_PER_LANE":4,"TGB_GI_POOL_SERVICE_SLOTS":64}]'
run pool_k2 k_gi_trace_pool '[{"TGB_GI_KERNEL":2,"TGB_GI_RAYS_PER_LANE":2,"TGB_GI_POOL_SERVICE_SLOTS":32}]'
ls -la gpurun_out/*.ncu-rep
