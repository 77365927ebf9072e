#!/bin/bash
# round 2, session 3: GI kernel 4 as the default. N = 1: whole GPU suite, bench line, ncu captures of the stage's three kernels. N > 1 (argument): multi-GPU tests + bench
N=${1:-1}
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  ( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r04m_pytest_gpu.log 2>&1; echo "exit $?" >> gpurun_out/r04m_pytest_gpu.log; tail -4 gpurun_out/r04m_pytest_gpu.log
  ( time timeout 600 python bench.py --steps 20 --warmup 5 ) > gpurun_out/r04m_bench_c2.json 2> gpurun_out/r04m_bench_c2.err; tail -c 200 gpurun_out/r04m_bench_c2.json
  for K in k_gi_trace_fast k_gi_trace_list k_shade; do
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K --launch-skip 3 -c 1 -o gpurun_out/r04m_$K -f python tools/sweep.py --workload c2 --frames 3 --configs '[{}]' > gpurun_out/r04m_$K.log 2>&1
  done
  K='regex:k_(clear|cull|sort|visibility|object_frames|shade|gi_|fast_|set_words|svo|resolve|present)'
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 400 --csv --log-file gpurun_out/r04m_launches_c2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --also none > gpurun_out/r04m_ncu_ll.log 2>&1
else
  ( time timeout 900 python -m pytest tests/test_multi_gpu.py -x -q -m gpu ) > gpurun_out/r04m_pytest_mgpu_n$N.log 2>&1; tail -3 gpurun_out/r04m_pytest_mgpu_n$N.log
  ( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 ) > gpurun_out/r04m_bench_n$N.json 2> gpurun_out/r04m_bench_n$N.err
  tail -c 600 gpurun_out/r04m_bench_n$N.json; tail -3 gpurun_out/r04m_bench_n$N.err
fi
