#!/bin/bash
# K1 with the cluster masks staged in shared memory by cp.async (TGB_K1_STAGE_MASKS=1): parity, timing on c2 / c2far, ncu
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_visibility_gpu.py -m gpu -x -q -k "staged or golden or config1" ) > gpurun_out/r03p_pytest_vis.log 2>&1; tail -3 gpurun_out/r03p_pytest_vis.log
CFG='[{},{"TGB_K1_STAGE_MASKS":1}]'
( timeout 300 python tools/sweep.py --workload c2 --frames 12 --what k1 --configs "$CFG" ) > gpurun_out/r03p_sweep_k1_stage.jsonl 2> gpurun_out/r03p_sweep_k1_stage.err
( timeout 300 python tools/sweep.py --workload c2far --frames 12 --what k1 --configs "$CFG" ) > gpurun_out/r03p_sweep_k1_stage_far.jsonl 2> gpurun_out/r03p_sweep_k1_stage_far.err
timeout 600 ncu --set full --clock-control none --import-source on -k k_visibility --launch-skip 3 -c 1 -o gpurun_out/r03p_k1_staged -f python tools/sweep.py --workload c2 --frames 3 --configs '[{"TGB_K1_STAGE_MASKS":1}]' > gpurun_out/r03p_k1_staged.log 2>&1
