#!/bin/bash
run() { r=$(env "$@" timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(l['config']['stage_ms']['shading_ms'],3), round(l['ms_per_step'],3))"); echo "$* -> shading_ms, frame_ms = $r"; }
run TGB_GI_POOL=0 TGB_GI_DDA_STEPS=16 TGB_GI_SERVICE_LANES=12
for t in 256 128; do for c in 4 5 8 10; do for k in 4 8 16; do
  if [ $t = 256 ] && [ $c -gt 5 ]; then continue; fi
  run TGB_GI_POOL=1 TGB_GI_POOL_THREADS=$t TGB_GI_POOL_CTAS_PER_SM=$c TGB_GI_DDA_STEPS=$k
done; done; done
