#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_procedural_gpu.py tests/test_shading_gpu.py -m gpu -x -q > gpurun_out/pytest_a.log 2>&1; echo "exit $?" >> gpurun_out/pytest_a.log; tail -3 gpurun_out/pytest_a.log
timeout 900 python -m pytest tests/test_visibility_gpu.py -m gpu -x -q -k config5 > gpurun_out/pytest_c5.log 2>&1; echo "exit $?" >> gpurun_out/pytest_c5.log; tail -5 gpurun_out/pytest_c5.log
timeout 300 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -c 300 gpurun_out/bench_c2.json
timeout 600 python bench.py --workload c5 > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; tail -c 300 gpurun_out/bench_c5.json; tail -3 gpurun_out/bench_c5.err
K='regex:k_(clear|cull|sort|visibility|object_frames|shade|gi_|set_words|svo|resolve)'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 400 --csv --log-file gpurun_out/launches_c2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_ll.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_visibility --launch-skip 3 -c 1 -o gpurun_out/r01m_k1 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_k1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_gi_trace_flat --launch-skip 3 -c 1 -o gpurun_out/r01m_k3b -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_k3b.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_shade --launch-skip 3 -c 1 -o gpurun_out/r01m_k3a -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_k3a.log 2>&1
ls -la gpurun_out
