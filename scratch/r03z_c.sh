#!/bin/bash
# after making list mode a template parameter of the pool kernel: shading tests, stage time, refreshed bench line and k3b capture
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_shading_gpu.py -m gpu -x -q ) > gpurun_out/r03z_pytest_shading.log 2>&1; tail -3 gpurun_out/r03z_pytest_shading.log
( timeout 600 python tools/sweep.py --workload c2 --frames 12 --configs '[{},{"TGB_GI_KERNEL":3}]' ) > gpurun_out/r03z_sweep_default.jsonl 2> gpurun_out/r03z_sweep_default.err
( time timeout 600 python bench.py --steps 20 --warmup 5 ) > gpurun_out/r03z_bench_c2.json 2> gpurun_out/r03z_bench_c2.err; tail -c 200 gpurun_out/r03z_bench_c2.json
timeout 600 ncu --set full --clock-control none --import-source on -k k_gi_trace_pool --launch-skip 3 -c 1 -o gpurun_out/r03z_k3b -f python tools/sweep.py --workload c2 --frames 3 --configs '[{}]' > gpurun_out/r03z_k3b.log 2>&1
K='regex:k_(clear|cull|sort|visibility|object_frames|shade|gi_|set_words|svo|resolve|present)'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 400 --csv --log-file gpurun_out/r03z_launches_c2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --also none > gpurun_out/r03z_ncu_ll.log 2>&1
