#!/bin/bash
# round 2, session 3: GI kernel 4 with runs of free voxels along the dominant axis; careful pass = the fast kernel once more over the hand-overs (cube check), then the exact list kernel
mkdir -p gpurun_out
( time TGB_GI_KERNEL=4 timeout 900 python -m pytest tests/test_shading_gpu.py -m gpu -x -q ) > gpurun_out/r04f_pytest_k4.log 2>&1; echo "exit $?" >> gpurun_out/r04f_pytest_k4.log; tail -5 gpurun_out/r04f_pytest_k4.log
CFG='[{"TGB_GI_KERNEL":2},{"TGB_GI_KERNEL":4},{"TGB_GI_KERNEL":4,"TGB_GI_FAST_CAREFUL":0},{"TGB_GI_KERNEL":4,"TGB_GI_LIST_KERNEL":0},{"TGB_GI_KERNEL":4,"TGB_GI_FAST_CAREFUL_CTAS_PER_SM":1},{"TGB_GI_KERNEL":4,"TGB_GI_FAST_CAREFUL_CTAS_PER_SM":8},{"TGB_GI_KERNEL":4,"TGB_GI_FAST_CAREFUL_SERVICE_LANES":8},{"TGB_GI_KERNEL":4,"TGB_GI_FAST_STEPS":4},{"TGB_GI_KERNEL":4,"TGB_GI_FAST_STEPS":6},{"TGB_GI_KERNEL":4,"TGB_GI_FAST_DELTA_PERCENT":25},{"TGB_GI_KERNEL":4,"TGB_GI_FAST_MAX_STEPS":64,"TGB_GI_FAST_MAX_STEPS_UNCERTAIN":16},{"TGB_GI_KERNEL":4,"TGB_GI_FAST_MAX_STEPS":1024,"TGB_GI_FAST_MAX_STEPS_UNCERTAIN":1024}]'
( time timeout 600 python tools/sweep.py --workload c2 --frames 10 --configs "$CFG" ) > gpurun_out/r04f_sweep_full.jsonl 2> gpurun_out/r04f_sweep_full.err
tail -2 gpurun_out/r04f_sweep_full.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_(gi_|fast_|shade|svo_flatten)' -c 40 --csv --log-file gpurun_out/r04f_ll.csv python tools/sweep.py --workload c2 --frames 3 --configs '[{"TGB_GI_KERNEL":4}]' > gpurun_out/r04f_ll.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gi_trace_fast --launch-skip 6 -c 2 -o gpurun_out/r04f_k_gi_trace_fast -f python tools/sweep.py --workload c2 --frames 3 --configs '[{"TGB_GI_KERNEL":4}]' > gpurun_out/r04f_k_gi_trace_fast.log 2>&1
