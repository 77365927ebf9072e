#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02o_pytest_gpu.log 2>&1; echo "exit $?" >> gpurun_out/r02o_pytest_gpu.log; tail -6 gpurun_out/r02o_pytest_gpu.log
for w in c4 c4m8; do
  ( timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --also none ) > gpurun_out/r02o_bench_$w.json 2> gpurun_out/r02o_bench_$w.err
done
python - <<'PY'
import json
for w in ("c4","c4m8"):
    r=json.loads(open(f"gpurun_out/r02o_bench_{w}.json").read().strip().splitlines()[-1])
    print(w, round(r["ms_per_step"],3), 'e2e', round(r['e2e']['ms_per_step'],3), round(r['e2e']['presented_bgra8']['ms_per_step'],3), {k:round(v,3) for k,v in r["config"]["stage_ms"].items()})
PY
