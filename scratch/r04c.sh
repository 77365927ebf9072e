#!/bin/bash
# round 2, session 3: GI kernel 4 with one margin per axis, CTA-aggregated queue append in k_shade; margin evidence on the device (DELTA at 25 % / 10 %)
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_shading_gpu.py -m gpu -x -q -k "certified or config3" ) > gpurun_out/r04c_pytest.log 2>&1; echo "exit $?" >> gpurun_out/r04c_pytest.log; tail -5 gpurun_out/r04c_pytest.log
CFG='[{"TGB_GI_KERNEL":2},{"TGB_GI_KERNEL":4,"TGB_GI_SHADE_STEPS":0,"TGB_GI_LIST_KERNEL":0},{"TGB_GI_KERNEL":4,"TGB_GI_SHADE_STEPS":0},{"TGB_GI_KERNEL":4},{"TGB_GI_KERNEL":4,"TGB_GI_LIST_KERNEL":0},{"TGB_GI_KERNEL":4,"TGB_GI_LIST_RAYS":8},{"TGB_GI_KERNEL":4,"TGB_GI_LIST_RAYS":16},{"TGB_GI_KERNEL":4,"TGB_GI_LIST_RAYS":8,"TGB_GI_LIST_DDA_STEPS":1024,"TGB_GI_LIST_TREE_REPS":64},{"TGB_GI_KERNEL":4,"TGB_GI_LIST_RAYS":2,"TGB_GI_LIST_DDA_STEPS":1024,"TGB_GI_LIST_TREE_REPS":64},{"TGB_GI_KERNEL":4,"TGB_GI_FAST_STEPS":4},{"TGB_GI_KERNEL":4,"TGB_GI_FAST_STEPS":16},{"TGB_GI_KERNEL":4,"TGB_GI_SHADE_STEPS":2},{"TGB_GI_KERNEL":4,"TGB_GI_FAST_DELTA_PERCENT":25},{"TGB_GI_KERNEL":4,"TGB_GI_FAST_DELTA_PERCENT":10},{"TGB_GI_KERNEL":4,"TGB_GI_FAST_SERVICE_LANES":4},{"TGB_GI_KERNEL":4,"TGB_GI_FAST_SERVICE_LANES":12}]'
( time timeout 600 python tools/sweep.py --workload c2 --frames 10 --configs "$CFG" ) > gpurun_out/r04c_sweep_full.jsonl 2> gpurun_out/r04c_sweep_full.err
tail -2 gpurun_out/r04c_sweep_full.err
CFG2='[{"TGB_GI_KERNEL":2},{"TGB_GI_KERNEL":4,"TGB_GI_SHADE_STEPS":0},{"TGB_GI_KERNEL":4},{"TGB_GI_KERNEL":4,"TGB_GI_LIST_KERNEL":0}]'
( time timeout 300 python tools/sweep.py --workload c2 --frames 10 --rows 272 --row0 944 --configs "$CFG2" ) > gpurun_out/r04c_sweep_tile.jsonl 2> gpurun_out/r04c_sweep_tile.err
( time timeout 300 python tools/sweep.py --workload c2far --frames 8 --configs '[{"TGB_GI_KERNEL":2},{"TGB_GI_KERNEL":4},{"TGB_GI_KERNEL":4,"TGB_GI_FAST_DELTA_PERCENT":25}]' ) > gpurun_out/r04c_sweep_far.jsonl 2> gpurun_out/r04c_sweep_far.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_(gi_|fast_|shade|svo_flatten)' -c 40 --csv --log-file gpurun_out/r04c_ll.csv python tools/sweep.py --workload c2 --frames 3 --configs '[{"TGB_GI_KERNEL":4}]' > gpurun_out/r04c_ll.log 2>&1
for K in k_gi_trace_fast k_gi_trace_list k_shade; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K --launch-skip 3 -c 1 -o gpurun_out/r04c_$K -f python tools/sweep.py --workload c2 --frames 3 --configs '[{"TGB_GI_KERNEL":4}]' > gpurun_out/r04c_$K.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_shade --launch-skip 3 -c 1 -o gpurun_out/r04c_k_shade_k2 -f python tools/sweep.py --workload c2 --frames 3 --configs '[{"TGB_GI_KERNEL":2}]' > gpurun_out/r04c_k_shade_k2.log 2>&1
ls -la gpurun_out/ | grep r04c | wc -l
