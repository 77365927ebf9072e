#!/bin/bash
mkdir -p gpurun_out
run() { r=$(env "$@" timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(l['config']['stage_ms']['shading_ms'],3), round(l['ms_per_step'],3))"); echo "$* -> shading_ms, frame_ms = $r" | tee -a gpurun_out/sweep_wpool.log; }
run TGB_GI_WPOOL=1 TGB_GI_WPOOL_DDA_STEPS=4
run TGB_GI_WPOOL=1 TGB_GI_WPOOL_DDA_STEPS=4 TGB_GI_WPOOL_CTAS_PER_SM=10
run TGB_GI_WPOOL=1 TGB_GI_WPOOL_DDA_STEPS=8 TGB_GI_WPOOL_CTAS_PER_SM=10
run TGB_GI_WPOOL=1 TGB_GI_WPOOL_SLOTS=128 TGB_GI_WPOOL_DDA_STEPS=8 TGB_GI_WPOOL_CTAS_PER_SM=5
run TGB_GI_WPOOL=1 TGB_GI_WPOOL_SLOTS=128 TGB_GI_WPOOL_DDA_STEPS=4 TGB_GI_WPOOL_CTAS_PER_SM=5
