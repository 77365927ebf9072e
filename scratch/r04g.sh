#!/bin/bash
# round 2, session 3: exact list kernel = one warp per ray, leaf blocks staged in shared memory by the warp; careful pass off
mkdir -p gpurun_out
( time TGB_GI_KERNEL=4 timeout 900 python -m pytest tests/test_shading_gpu.py -m gpu -x -q ) > gpurun_out/r04g_pytest_k4.log 2>&1; echo "exit $?" >> gpurun_out/r04g_pytest_k4.log; tail -5 gpurun_out/r04g_pytest_k4.log
CFG='[{"TGB_GI_KERNEL":2},{"TGB_GI_KERNEL":4},{"TGB_GI_KERNEL":4,"TGB_GI_LIST_KERNEL":0},{"TGB_GI_KERNEL":4,"TGB_GI_FAST_CAREFUL":1},{"TGB_GI_KERNEL":4,"TGB_GI_LIST_CTAS_PER_SM":16},{"TGB_GI_KERNEL":4,"TGB_GI_LIST_TREE_REPS":4,"TGB_GI_LIST_DDA_STEPS":64},{"TGB_GI_KERNEL":4,"TGB_GI_FAST_STEPS":4},{"TGB_GI_KERNEL":4,"TGB_GI_FAST_STEPS":6},{"TGB_GI_KERNEL":4,"TGB_GI_FAST_DELTA_PERCENT":25},{"TGB_GI_KERNEL":4,"TGB_GI_FAST_MAX_STEPS":64,"TGB_GI_FAST_MAX_STEPS_UNCERTAIN":16},{"TGB_GI_KERNEL":4,"TGB_GI_FAST_SERVICE_LANES":12},{"TGB_GI_KERNEL":4,"TGB_GI_FAST_SERVICE_LANES":6}]'
( time timeout 600 python tools/sweep.py --workload c2 --frames 10 --configs "$CFG" ) > gpurun_out/r04g_sweep_full.jsonl 2> gpurun_out/r04g_sweep_full.err
tail -2 gpurun_out/r04g_sweep_full.err
CFG2='[{"TGB_GI_KERNEL":2},{"TGB_GI_KERNEL":4},{"TGB_GI_KERNEL":4,"TGB_GI_LIST_KERNEL":0}]'
( time timeout 300 python tools/sweep.py --workload c2 --frames 10 --rows 272 --row0 944 --configs "$CFG2" ) > gpurun_out/r04g_sweep_tile.jsonl 2> gpurun_out/r04g_sweep_tile.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_(gi_|fast_|shade|svo_flatten)' -c 40 --csv --log-file gpurun_out/r04g_ll.csv python tools/sweep.py --workload c2 --frames 3 --configs '[{"TGB_GI_KERNEL":4}]' > gpurun_out/r04g_ll.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gi_trace_list --launch-skip 3 -c 1 -o gpurun_out/r04g_k_gi_trace_list -f python tools/sweep.py --workload c2 --frames 3 --configs '[{"TGB_GI_KERNEL":4}]' > gpurun_out/r04g_k_gi_trace_list.log 2>&1
