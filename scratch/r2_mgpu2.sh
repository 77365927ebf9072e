#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_shading_gpu.py tests/test_multi_gpu.py -m gpu -x -q > gpurun_out/pytest_mgpu.log 2>&1; echo "exit $?" >> gpurun_out/pytest_mgpu.log; tail -5 gpurun_out/pytest_mgpu.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 2 --steps 20 --warmup 3 --merge peer > gpurun_out/bench_c2_n2_peer.json 2> gpurun_out/bench_c2_n2_peer.err
python - <<'PY'
import json
l=json.loads(open("gpurun_out/bench_c2_n2_peer.json").read().strip().splitlines()[-1])
print(round(l["value"]), round(l["ms_per_step"],3), {k:round(v,3) for k,v in l["config"]["stage_ms"].items()})
PY
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
python - <<'PY'
import json
l=json.loads(open("gpurun_out/bench_c2.json").read().strip().splitlines()[-1])
print(round(l["value"]), round(l["ms_per_step"],3), {k:round(v,3) for k,v in l["config"]["stage_ms"].items()}, "e2e", round(l["e2e"]["value"]))
PY
