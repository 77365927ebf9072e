#!/bin/bash
# round 2, session 3: list kernel with the branch form of the leaf DDA
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_shading_gpu.py -m gpu -x -q ) > gpurun_out/r04p_pytest.log 2>&1; echo "exit $?" >> gpurun_out/r04p_pytest.log; tail -4 gpurun_out/r04p_pytest.log
CFG='[{"TGB_GI_KERNEL":2},{},{"TGB_GI_LIST_TMA":0},{},{"TGB_GI_LIST_CTAS_PER_SM":16}]'
( time timeout 600 python tools/sweep.py --workload c2 --frames 12 --configs "$CFG" ) > gpurun_out/r04p_sweep_full.jsonl 2> gpurun_out/r04p_sweep_full.err
( time timeout 300 python tools/sweep.py --workload c2 --frames 10 --rows 272 --row0 944 --configs '[{"TGB_GI_KERNEL":2},{},{}]' ) > gpurun_out/r04p_sweep_tile.jsonl 2> gpurun_out/r04p_sweep_tile.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gi_trace_list --launch-skip 3 -c 1 -o gpurun_out/r04p_k_gi_trace_list -f python tools/sweep.py --workload c2 --frames 3 --configs '[{}]' > gpurun_out/r04p_k_gi_trace_list.log 2>&1
