#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "exit $?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -c 200 gpurun_out/bench_c2.json
timeout 400 python bench.py --workload c4 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; tail -c 200 gpurun_out/bench_c4.json
timeout 600 python bench.py --workload c5 > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; tail -c 200 gpurun_out/bench_c5.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 200 gpurun_out/bench_ref.json
K='regex:k_(clear|cull|sort|visibility|object_frames|shade|gi_|set_words|svo|resolve|present)'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 400 --csv --log-file gpurun_out/launches_c2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_ll.log 2>&1
for spec in "k1:k_visibility:3" "k3a:k_shade:3" "k3b:k_gi_trace_flat:3" "k2:k_svo_fill_leaves:1"; do
  IFS=: read tag kern skip <<< "$spec"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$kern --launch-skip $skip -c 1 -o gpurun_out/r01p_$tag -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_$tag.log 2>&1
done
ls -la gpurun_out | head -40
