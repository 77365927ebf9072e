#!/bin/bash
# fast GI walk with batch set-up, fair-share exact pass: shading tests, full frame and tile-sized batch, launch list
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_shading_gpu.py -m gpu -x -q ) > gpurun_out/r03i_pytest_shading.log 2>&1; echo "exit $?" >> gpurun_out/r03i_pytest_shading.log; tail -4 gpurun_out/r03i_pytest_shading.log
CFG='[{"TGB_GI_KERNEL":2},{"TGB_GI_KERNEL":3},{"TGB_GI_KERNEL":3,"TGB_GI_FAST_STEPS":8},{"TGB_GI_KERNEL":3,"TGB_GI_FAST_STEPS":8,"TGB_GI_FAST_SERVICE_LANES":12},{"TGB_GI_KERNEL":3,"TGB_GI_FAST_STEPS":8,"TGB_GI_FAST_SERVICE_LANES":4},{"TGB_GI_KERNEL":3,"TGB_GI_FAST_STEPS":8,"TGB_GI_RAYS_PER_LANE":1,"TGB_GI_POOL_CTAS_PER_SM":8},{"TGB_GI_KERNEL":3,"TGB_GI_FAST_STEPS":8,"TGB_GI_RAYS_PER_LANE":1,"TGB_GI_POOL_CTAS_PER_SM":8,"TGB_GI_POOL_SERVICE_SLOTS":4},{"TGB_GI_KERNEL":3,"TGB_GI_FAST_STEPS":8,"TGB_GI_RAYS_PER_LANE":2,"TGB_GI_POOL_CTAS_PER_SM":8}]'
( time timeout 600 python tools/sweep.py --workload c2 --frames 10 --rows 272 --row0 1088 --configs "$CFG" ) > gpurun_out/r03i_sweep_tile.jsonl 2> gpurun_out/r03i_sweep_tile.err
( time timeout 600 python tools/sweep.py --workload c2 --frames 10 --configs "$CFG" ) > gpurun_out/r03i_sweep_full.jsonl 2> gpurun_out/r03i_sweep_full.err
tail -2 gpurun_out/r03i_sweep_full.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_gi_ -c 40 --csv --log-file gpurun_out/r03i_ll.csv python tools/sweep.py --workload c2 --frames 3 --configs '[{"TGB_GI_FAST_STEPS":8,"TGB_GI_RAYS_PER_LANE":1,"TGB_GI_POOL_CTAS_PER_SM":8}]' > gpurun_out/r03i_ll.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gi_trace_fast --launch-skip 3 -c 1 -o gpurun_out/r03i_k3fast -f python tools/sweep.py --workload c2 --frames 3 --configs '[{"TGB_GI_FAST_STEPS":8}]' > gpurun_out/r03i_k3fast.log 2>&1
