#!/bin/bash
# round 2, session 3: TGB_GI_KERNEL=4 with k_shade taking the first certified steps, k_gi_trace_list for the hand-overs, unified cell decode
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_shading_gpu.py -m gpu -x -q -k "certified or config3" ) > gpurun_out/r04b_pytest.log 2>&1; echo "exit $?" >> gpurun_out/r04b_pytest.log; tail -5 gpurun_out/r04b_pytest.log
CFG='[{"TGB_GI_KERNEL":2},{"TGB_GI_KERNEL":4,"TGB_GI_SHADE_STEPS":0,"TGB_GI_LIST_KERNEL":0},{"TGB_GI_KERNEL":4,"TGB_GI_SHADE_STEPS":0},{"TGB_GI_KERNEL":4},{"TGB_GI_KERNEL":4,"TGB_GI_SHADE_STEPS":2},{"TGB_GI_KERNEL":4,"TGB_GI_SHADE_STEPS":4},{"TGB_GI_KERNEL":4,"TGB_GI_SHADE_STEPS":8},{"TGB_GI_KERNEL":4,"TGB_GI_FAST_STEPS":2},{"TGB_GI_KERNEL":4,"TGB_GI_FAST_STEPS":4},{"TGB_GI_KERNEL":4,"TGB_GI_FAST_STEPS":16},{"TGB_GI_KERNEL":4,"TGB_GI_FAST_SERVICE_LANES":4},{"TGB_GI_KERNEL":4,"TGB_GI_FAST_SERVICE_LANES":16},{"TGB_GI_KERNEL":4,"TGB_GI_LIST_RAYS":1},{"TGB_GI_KERNEL":4,"TGB_GI_LIST_RAYS":2},{"TGB_GI_KERNEL":4,"TGB_GI_LIST_RAYS":8},{"TGB_GI_KERNEL":4,"TGB_GI_LIST_RAYS":32},{"TGB_GI_KERNEL":4,"TGB_GI_LIST_DDA_STEPS":1024,"TGB_GI_LIST_TREE_REPS":64},{"TGB_GI_KERNEL":4,"TGB_GI_FAST_CTAS_PER_SM":6},{"TGB_GI_KERNEL":4,"TGB_GI_FAST_CTAS_PER_SM":10}]'
( time timeout 600 python tools/sweep.py --workload c2 --frames 10 --configs "$CFG" ) > gpurun_out/r04b_sweep_full.jsonl 2> gpurun_out/r04b_sweep_full.err
tail -2 gpurun_out/r04b_sweep_full.err
CFG2='[{"TGB_GI_KERNEL":2},{"TGB_GI_KERNEL":4,"TGB_GI_SHADE_STEPS":0},{"TGB_GI_KERNEL":4},{"TGB_GI_KERNEL":4,"TGB_GI_SHADE_STEPS":2},{"TGB_GI_KERNEL":4,"TGB_GI_FAST_STEPS":4}]'
( time timeout 300 python tools/sweep.py --workload c2 --frames 10 --rows 272 --row0 944 --configs "$CFG2" ) > gpurun_out/r04b_sweep_tile.jsonl 2> gpurun_out/r04b_sweep_tile.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_(gi_|fast_|shade|svo_flatten)' -c 40 --csv --log-file gpurun_out/r04b_ll.csv python tools/sweep.py --workload c2 --frames 3 --configs '[{"TGB_GI_KERNEL":4}]' > gpurun_out/r04b_ll.log 2>&1
for K in k_gi_trace_fast k_gi_trace_list k_shade; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K --launch-skip 3 -c 1 -o gpurun_out/r04b_$K -f python tools/sweep.py --workload c2 --frames 3 --configs '[{"TGB_GI_KERNEL":4}]' > gpurun_out/r04b_$K.log 2>&1
done
ls -la gpurun_out/ | grep r04b
