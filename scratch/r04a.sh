#!/bin/bash
# round 2, session 3: the certified fast walk over the coarser tiling (TGB_GI_KERNEL=4): parity tests, sweep against kernels 2 / 3, launch list, ncu capture
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_shading_gpu.py -m gpu -x -q -k "certified or config3" ) > gpurun_out/r04a_pytest.log 2>&1; echo "exit $?" >> gpurun_out/r04a_pytest.log; tail -5 gpurun_out/r04a_pytest.log
CFG='[{"TGB_GI_KERNEL":2},{"TGB_GI_KERNEL":3},{"TGB_GI_KERNEL":4},{"TGB_GI_KERNEL":4,"TGB_GI_FAST_STEPS":2},{"TGB_GI_KERNEL":4,"TGB_GI_FAST_STEPS":4},{"TGB_GI_KERNEL":4,"TGB_GI_FAST_STEPS":16},{"TGB_GI_KERNEL":4,"TGB_GI_FAST_SERVICE_LANES":4},{"TGB_GI_KERNEL":4,"TGB_GI_FAST_SERVICE_LANES":12},{"TGB_GI_KERNEL":4,"TGB_GI_FAST_SERVICE_LANES":16},{"TGB_GI_KERNEL":4,"TGB_GI_FAST_SERVICE_LANES":24},{"TGB_GI_KERNEL":4,"TGB_GI_FAST_CTAS_PER_SM":4},{"TGB_GI_KERNEL":4,"TGB_GI_FAST_CTAS_PER_SM":6},{"TGB_GI_KERNEL":4,"TGB_GI_FAST_STEPS":4,"TGB_GI_FAST_SERVICE_LANES":16},{"TGB_GI_KERNEL":4,"TGB_GI_RAYS_PER_LANE":1,"TGB_GI_POOL_CTAS_PER_SM":4}]'
( time timeout 600 python tools/sweep.py --workload c2 --frames 10 --configs "$CFG" ) > gpurun_out/r04a_sweep_full.jsonl 2> gpurun_out/r04a_sweep_full.err
tail -2 gpurun_out/r04a_sweep_full.err
CFG2='[{"TGB_GI_KERNEL":2},{"TGB_GI_KERNEL":3},{"TGB_GI_KERNEL":4},{"TGB_GI_KERNEL":4,"TGB_GI_FAST_STEPS":4},{"TGB_GI_KERNEL":4,"TGB_GI_FAST_SERVICE_LANES":16}]'
( time timeout 300 python tools/sweep.py --workload c2 --frames 10 --rows 272 --row0 944 --configs "$CFG2" ) > gpurun_out/r04a_sweep_tile.jsonl 2> gpurun_out/r04a_sweep_tile.err
( time timeout 300 python tools/sweep.py --workload c2far --frames 8 --configs '[{"TGB_GI_KERNEL":2},{"TGB_GI_KERNEL":4}]' ) > gpurun_out/r04a_sweep_far.jsonl 2> gpurun_out/r04a_sweep_far.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_(gi_|fast_|shade|svo_flatten)' -c 40 --csv --log-file gpurun_out/r04a_ll.csv python tools/sweep.py --workload c2 --frames 3 --configs '[{"TGB_GI_KERNEL":4}]' > gpurun_out/r04a_ll.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gi_trace_fast --launch-skip 3 -c 1 -o gpurun_out/r04a_k3tiled -f python tools/sweep.py --workload c2 --frames 3 --configs '[{"TGB_GI_KERNEL":4}]' > gpurun_out/r04a_k3tiled.log 2>&1
ls -la gpurun_out/ | tail -12
