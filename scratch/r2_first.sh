#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 600 gpurun_out/bench_default.json
run() { r=$(env "$@" timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(l['config']['stage_ms']['shading_ms'],3), round(l['config']['stage_ms']['visibility_ms'],3), round(l['ms_per_step'],3))"); echo "$* -> shading_ms, vis_ms, frame_ms = $r" | tee -a gpurun_out/sweep.log; }
run TGB_GI_POOL=0
run TGB_GI_POOL=0 TGB_GI_DDA_STEPS=16 TGB_GI_SERVICE_LANES=12
run TGB_GI_POOL=1
run TGB_GI_POOL=1 TGB_GI_POOL_THREADS=128 TGB_GI_POOL_CTAS_PER_SM=8
run TGB_GI_POOL=1 TGB_GI_POOL_THREADS=128 TGB_GI_POOL_CTAS_PER_SM=10 TGB_GI_DDA_STEPS=16
run TGB_GI_POOL=1 TGB_GI_POOL_THREADS=256 TGB_GI_POOL_CTAS_PER_SM=4 TGB_GI_DDA_STEPS=16
