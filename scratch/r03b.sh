#!/bin/bash
# round 2, session 2: unified-step fast GI walk (tgb_gi_fast.cu): shading tests, knob sweep, ncu capture
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_shading_gpu.py -m gpu -x -q ) > gpurun_out/r03c_pytest_shading.log 2>&1; echo "exit $?" >> gpurun_out/r03c_pytest_shading.log; tail -5 gpurun_out/r03c_pytest_shading.log
( time timeout 600 python tools/sweep.py --workload c2 --frames 10 --what fast ) > gpurun_out/r03c_sweep_fast.jsonl 2> gpurun_out/r03c_sweep_fast.err; tail -3 gpurun_out/r03c_sweep_fast.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gi_trace_fast --launch-skip 3 -c 1 -o gpurun_out/r03c_k3fast -f python tools/sweep.py --workload c2 --frames 3 --configs '[{}]' > gpurun_out/r03c_k3fast.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gi_trace_pool --launch-skip 3 -c 1 -o gpurun_out/r03c_k3exact -f python tools/sweep.py --workload c2 --frames 3 --configs '[{}]' > gpurun_out/r03c_k3exact.log 2>&1
ls -la gpurun_out/r03c_*
