#!/bin/bash
# round 2, session 2: the record of the final tree on one GPU: GPU suite, bench lines, ncu launch list and --set full captures
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r03z_pytest_gpu.log 2>&1; echo "exit $?" >> gpurun_out/r03z_pytest_gpu.log; tail -5 gpurun_out/r03z_pytest_gpu.log
( time timeout 600 python bench.py --steps 20 --warmup 5 ) > gpurun_out/r03z_bench_c2.json 2> gpurun_out/r03z_bench_c2.err; tail -c 300 gpurun_out/r03z_bench_c2.json
( time timeout 300 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/r03z_bench_ref.json 2> gpurun_out/r03z_bench_ref.err; tail -c 300 gpurun_out/r03z_bench_ref.json
K='regex:k_(clear|cull|sort|visibility|object_frames|shade|gi_|set_words|svo|resolve|present)'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 400 --csv --log-file gpurun_out/r03z_launches_c2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --also none > gpurun_out/r03z_ncu_ll.log 2>&1
run() { # tag, kernel regex, workload, config json
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 --launch-skip 3 -c 1 -o gpurun_out/r03z_$1 -f python tools/sweep.py --workload $3 --frames 3 --configs "$4" > gpurun_out/r03z_$1.log 2>&1
}
run k1 'k_visibility<' c2 '[{}]'
run k3a k_shade c2 '[{}]'
run k3b k_gi_trace_pool c2 '[{}]'
run k1_far 'k_visibility<' c2far '[{}]'
run k2 k_svo_fill_leaves c2 '[{}]'
ls -la gpurun_out/r03z_*.ncu-rep
