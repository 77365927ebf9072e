#!/bin/bash
# sweep of the GI scheduling knobs: prints the shading stage time per setting
for k in 4 8 16 64; do for s in 4 8 12; do for b in 0 4 8; do
  r=$(TGB_GI_DDA_STEPS=$k TGB_GI_SERVICE_LANES=$s TGB_GI_DDA_BIAS=$b timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(l['config']['stage_ms']['shading_ms'],3))")
  echo "steps=$k service=$s bias=$b shading_ms=$r"
done; done; done
