#!/bin/bash
mkdir -p gpurun_out
run() { r=$(env "$@" timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(l['config']['stage_ms']['shading_ms'],3), round(l['config']['stage_ms']['visibility_ms'],3), round(l['ms_per_step'],3))"); echo "$* -> shading_ms, vis_ms, frame_ms = $r" | tee -a gpurun_out/sweep_q.log; }
run TGB_GI_DDA_QUORUM=0
run TGB_GI_DDA_QUORUM=1
run TGB_GI_DDA_QUORUM=2
run TGB_GI_DDA_QUORUM=3
run TGB_GI_DDA_QUORUM=1 TGB_GI_DDA_STEPS=64
run TGB_GI_DDA_QUORUM=2 TGB_GI_DDA_STEPS=64
run TGB_GI_DDA_QUORUM=2 TGB_GI_DDA_STEPS=64 TGB_GI_SERVICE_LANES=8
run TGB_GI_DDA_QUORUM=2 TGB_GI_DDA_STEPS=64 TGB_GI_SERVICE_LANES=16
timeout 600 python -m pytest tests/test_shading_gpu.py -m gpu -x -q 2>&1 | tail -2
