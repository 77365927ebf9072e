#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_all.log 2>&1; echo "exit $?" >> gpurun_out/pytest_all.log; tail -12 gpurun_out/pytest_all.log
for m in peer nccl; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 2 --steps 20 --warmup 3 --merge $m > gpurun_out/bench_c2_n2_$m.json 2> gpurun_out/bench_c2_n2_$m.err
grep -v "^\*\*\*\|OMP_NUM" gpurun_out/bench_c2_n2_$m.err | tail -5
python - <<PY
import json
l=json.loads(open("gpurun_out/bench_c2_n2_$m.json").read().strip().splitlines()[-1])
print("$m", round(l["value"]), round(l["ms_per_step"],3), {k:round(v,3) for k,v in l["config"]["stage_ms"].items()}, "e2e", round(l["e2e"]["value"]), "present", round(l["e2e"]["presented_bgra8"]["value"]))
PY
done
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
python - <<'PY'
import json
l=json.loads(open("gpurun_out/bench_c2.json").read().strip().splitlines()[-1])
print("n1", round(l["value"]), round(l["ms_per_step"],3), {k:round(v,3) for k,v in l["config"]["stage_ms"].items()}, "e2e", round(l["e2e"]["value"]))
PY
