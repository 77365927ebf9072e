#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_svo_gpu.py -m gpu -x -q -s > gpurun_out/pytest_svo.log 2>&1; echo "exit $?" >> gpurun_out/pytest_svo.log; tail -8 gpurun_out/pytest_svo.log
timeout 400 python bench.py --workload c4 --no-cpu-baseline > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; tail -3 gpurun_out/bench_c4.err
python - <<'PY'
import json
l=json.loads(open("gpurun_out/bench_c4.json").read().strip().splitlines()[-1])
print("c4", round(l["value"]), round(l["ms_per_step"],3), l["config"]["stage_ms"], l["config"]["svo_build_ms"])
PY
