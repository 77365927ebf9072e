#!/bin/bash
# round 2, session 3: GI kernel 4 with the cube check of near-edge steps and shallow directions taken; whole GPU suite and bench with kernel 4 selected
mkdir -p gpurun_out
( time TGB_GI_KERNEL=4 timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r04d_pytest_k4.log 2>&1; echo "exit $?" >> gpurun_out/r04d_pytest_k4.log; tail -5 gpurun_out/r04d_pytest_k4.log
CFG='[{"TGB_GI_KERNEL":2},{"TGB_GI_KERNEL":4},{"TGB_GI_KERNEL":4,"TGB_GI_LIST_KERNEL":0},{"TGB_GI_KERNEL":4,"TGB_GI_LIST_RAYS":2},{"TGB_GI_KERNEL":4,"TGB_GI_LIST_RAYS":4},{"TGB_GI_KERNEL":4,"TGB_GI_LIST_CTAS_PER_SM":8},{"TGB_GI_KERNEL":4,"TGB_GI_FAST_STEPS":4},{"TGB_GI_KERNEL":4,"TGB_GI_FAST_STEPS":6},{"TGB_GI_KERNEL":4,"TGB_GI_FAST_DELTA_PERCENT":25},{"TGB_GI_KERNEL":4,"TGB_GI_FAST_DELTA_PERCENT":10},{"TGB_GI_KERNEL":4,"TGB_GI_FAST_SERVICE_LANES":6},{"TGB_GI_KERNEL":4,"TGB_GI_FAST_CTAS_PER_SM":7},{"TGB_GI_KERNEL":4,"TGB_GI_FAST_MAX_STEPS":1024,"TGB_GI_FAST_MAX_STEPS_UNCERTAIN":1024}]'
( time timeout 600 python tools/sweep.py --workload c2 --frames 10 --configs "$CFG" ) > gpurun_out/r04d_sweep_full.jsonl 2> gpurun_out/r04d_sweep_full.err
tail -2 gpurun_out/r04d_sweep_full.err
CFG2='[{"TGB_GI_KERNEL":2},{"TGB_GI_KERNEL":4},{"TGB_GI_KERNEL":4,"TGB_GI_LIST_KERNEL":0}]'
( time timeout 300 python tools/sweep.py --workload c2 --frames 10 --rows 272 --row0 944 --configs "$CFG2" ) > gpurun_out/r04d_sweep_tile.jsonl 2> gpurun_out/r04d_sweep_tile.err
( time timeout 300 python tools/sweep.py --workload c2far --frames 8 --configs '[{"TGB_GI_KERNEL":2},{"TGB_GI_KERNEL":4},{"TGB_GI_KERNEL":4,"TGB_GI_FAST_DELTA_PERCENT":25}]' ) > gpurun_out/r04d_sweep_far.jsonl 2> gpurun_out/r04d_sweep_far.err
( time timeout 300 python tools/sweep.py --workload c5 --frames 8 --configs '[{"TGB_GI_KERNEL":2},{"TGB_GI_KERNEL":4},{"TGB_GI_KERNEL":4,"TGB_GI_FAST_DELTA_PERCENT":25}]' ) > gpurun_out/r04d_sweep_c5.jsonl 2> gpurun_out/r04d_sweep_c5.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_(gi_|fast_|shade|svo_flatten)' -c 40 --csv --log-file gpurun_out/r04d_ll.csv python tools/sweep.py --workload c2 --frames 3 --configs '[{"TGB_GI_KERNEL":4}]' > gpurun_out/r04d_ll.log 2>&1
( time TGB_GI_KERNEL=4 timeout 600 python bench.py --steps 20 --warmup 5 ) > gpurun_out/r04d_bench_k4.json 2> gpurun_out/r04d_bench_k4.err; tail -c 300 gpurun_out/r04d_bench_k4.json
for K in k_gi_trace_fast k_gi_trace_list; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K --launch-skip 3 -c 1 -o gpurun_out/r04d_$K -f python tools/sweep.py --workload c2 --frames 3 --configs '[{"TGB_GI_KERNEL":4}]' > gpurun_out/r04d_$K.log 2>&1
done
ls gpurun_out | grep r04d | wc -l
