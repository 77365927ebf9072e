#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_shading_gpu.py tests/test_procedural_gpu.py -m gpu -x -q > gpurun_out/pytest_b.log 2>&1; echo "exit $?" >> gpurun_out/pytest_b.log; tail -6 gpurun_out/pytest_b.log
timeout 300 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -c 300 gpurun_out/bench_c2.json; tail -3 gpurun_out/bench_c2.err
timeout 400 python bench.py --workload c4 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; tail -c 300 gpurun_out/bench_c4.json; tail -3 gpurun_out/bench_c4.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 300 gpurun_out/bench_ref.json
