#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_visibility_gpu.py tests/test_multi_gpu.py -m gpu -x -q > gpurun_out/pytest_k1.log 2>&1; echo "exit $?" >> gpurun_out/pytest_k1.log; tail -4 gpurun_out/pytest_k1.log
run() { r=$(env "$@" timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(l['config']['stage_ms']['shading_ms'],3), round(l['config']['stage_ms']['visibility_ms'],3), round(l['ms_per_step'],3))"); echo "$* -> shading_ms, vis_ms, frame_ms = $r" | tee -a gpurun_out/sweep_k1.log; }
run TGB_K1_REGROUP=0 TGB_K1_MIN_CTAS=4
run TGB_K1_REGROUP=1 TGB_K1_MIN_CTAS=4
run TGB_K1_REGROUP=1 TGB_K1_MIN_CTAS=3
run TGB_K1_REGROUP=0 TGB_K1_MIN_CTAS=3
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_visibility --launch-skip 3 -c 1 -o gpurun_out/r01n_k1 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_k1.log 2>&1
