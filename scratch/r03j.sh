#!/bin/bash
# fast GI walk: step caps, launch list of the two passes
mkdir -p gpurun_out
CFG='[{"TGB_GI_KERNEL":2},{"TGB_GI_KERNEL":3},{"TGB_GI_KERNEL":3,"TGB_GI_FAST_MAX_STEPS_UNCERTAIN":32},{"TGB_GI_KERNEL":3,"TGB_GI_FAST_MAX_STEPS_UNCERTAIN":128},{"TGB_GI_KERNEL":3,"TGB_GI_FAST_MAX_STEPS_UNCERTAIN":256},{"TGB_GI_KERNEL":3,"TGB_GI_FAST_MAX_STEPS":128,"TGB_GI_FAST_MAX_STEPS_UNCERTAIN":32},{"TGB_GI_KERNEL":3,"TGB_GI_FAST_MAX_STEPS":512,"TGB_GI_FAST_MAX_STEPS_UNCERTAIN":128},{"TGB_GI_KERNEL":3,"TGB_GI_POOL_CTAS_PER_SM":4},{"TGB_GI_KERNEL":3,"TGB_GI_RAYS_PER_LANE":3}]'
( time timeout 600 python tools/sweep.py --workload c2 --frames 10 --configs "$CFG" ) > gpurun_out/r03j_sweep_full.jsonl 2> gpurun_out/r03j_sweep_full.err
tail -2 gpurun_out/r03j_sweep_full.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_gi_ -c 30 --csv --log-file gpurun_out/r03j_ll.csv python tools/sweep.py --workload c2 --frames 3 --configs '[{"TGB_GI_KERNEL":3}]' > gpurun_out/r03j_ll.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gi_trace_fast --launch-skip 3 -c 1 -o gpurun_out/r03j_k3fast -f python tools/sweep.py --workload c2 --frames 3 --configs '[{"TGB_GI_KERNEL":3}]' > gpurun_out/r03j_k3fast.log 2>&1
