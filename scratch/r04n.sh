#!/bin/bash
# round 2, session 3: a rank's tile-sized batch (272 rows): step caps, service threshold, phase length of the certified walk
mkdir -p gpurun_out
CFG='[{"TGB_GI_KERNEL":2},{},{"TGB_GI_FAST_MAX_STEPS":128,"TGB_GI_FAST_MAX_STEPS_UNCERTAIN":32},{"TGB_GI_FAST_MAX_STEPS":64,"TGB_GI_FAST_MAX_STEPS_UNCERTAIN":16},{"TGB_GI_FAST_MAX_STEPS":32,"TGB_GI_FAST_MAX_STEPS_UNCERTAIN":8},{"TGB_GI_FAST_SERVICE_LANES":2},{"TGB_GI_FAST_SERVICE_LANES":4},{"TGB_GI_FAST_STEPS":2},{"TGB_GI_FAST_STEPS":8},{"TGB_GI_FAST_CTAS_PER_SM":4},{"TGB_GI_SHADE_STEPS":2},{"TGB_GI_SHADE_STEPS":4},{"TGB_GI_FAST_MAX_STEPS":64,"TGB_GI_FAST_MAX_STEPS_UNCERTAIN":16,"TGB_GI_FAST_SERVICE_LANES":4},{}]'
for R0 in 944 400; do
( time timeout 300 python tools/sweep.py --workload c2 --frames 12 --rows 272 --row0 $R0 --configs "$CFG" ) > gpurun_out/r04n_sweep_tile_$R0.jsonl 2> gpurun_out/r04n_sweep_tile_$R0.err
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_(gi_|fast_|shade|object_frames|set_words)' -c 60 --csv --log-file gpurun_out/r04n_ll_tile.csv python tools/sweep.py --workload c2 --frames 3 --rows 272 --row0 944 --configs '[{}]' > gpurun_out/r04n_ll.log 2>&1
