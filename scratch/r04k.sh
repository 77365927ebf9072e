#!/bin/bash
# round 2, session 3: leaf blocks staged by a TMA bulk copy in k_gi_trace_list; k_shade<FAST> at 5 CTAs per SM
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_shading_gpu.py -m gpu -x -q ) > gpurun_out/r04k_pytest.log 2>&1; echo "exit $?" >> gpurun_out/r04k_pytest.log; tail -4 gpurun_out/r04k_pytest.log
CFG='[{"TGB_GI_KERNEL":2},{},{"TGB_GI_LIST_TMA":0},{"TGB_SHADE_MIN_CTAS":5},{"TGB_SHADE_MIN_CTAS":5,"TGB_GI_LIST_TMA":0},{"TGB_GI_FAST_SERVICE_LANES":6},{"TGB_SHADE_MIN_CTAS":5,"TGB_GI_FAST_SERVICE_LANES":6},{},{"TGB_GI_LIST_TMA":0},{"TGB_SHADE_MIN_CTAS":5}]'
( time timeout 600 python tools/sweep.py --workload c2 --frames 12 --configs "$CFG" ) > gpurun_out/r04k_sweep_full.jsonl 2> gpurun_out/r04k_sweep_full.err
tail -2 gpurun_out/r04k_sweep_full.err
CFG2='[{"TGB_GI_KERNEL":2},{},{"TGB_GI_LIST_TMA":0},{"TGB_SHADE_MIN_CTAS":5}]'
( time timeout 300 python tools/sweep.py --workload c2 --frames 10 --rows 272 --row0 944 --configs "$CFG2" ) > gpurun_out/r04k_sweep_tile.jsonl 2> gpurun_out/r04k_sweep_tile.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gi_trace_list --launch-skip 3 -c 1 -o gpurun_out/r04k_k_gi_trace_list -f python tools/sweep.py --workload c2 --frames 3 --configs '[{}]' > gpurun_out/r04k_k_gi_trace_list.log 2>&1
SEL="not config2 and not config3 and not config4 and not config5 and not full_size"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 99 --print-limit 20 python -m pytest tests/test_shading_gpu.py -m gpu -x -q -k "$SEL and (fast_walk or full_frame or golden)" > gpurun_out/r04k_memcheck.log 2>&1; echo "exit $?" >> gpurun_out/r04k_memcheck.log; tail -4 gpurun_out/r04k_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 99 --print-limit 20 python -m pytest tests/test_shading_gpu.py -m gpu -x -q -k "$SEL and fast_walk" > gpurun_out/r04k_racecheck.log 2>&1; echo "exit $?" >> gpurun_out/r04k_racecheck.log; tail -4 gpurun_out/r04k_racecheck.log
