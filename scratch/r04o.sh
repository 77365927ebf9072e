#!/bin/bash
# round 2, session 3: final tree. N = 1: GPU suite + bench line; N > 1: bench line
N=${1:-1}
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  ( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r04o_pytest_gpu.log 2>&1; echo "exit $?" >> gpurun_out/r04o_pytest_gpu.log; tail -4 gpurun_out/r04o_pytest_gpu.log
  ( time timeout 600 python bench.py --steps 20 --warmup 5 ) > gpurun_out/r04o_bench_c2.json 2> gpurun_out/r04o_bench_c2.err; tail -c 200 gpurun_out/r04o_bench_c2.json
  ( time timeout 300 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/r04o_bench_ref.json 2> gpurun_out/r04o_bench_ref.err; tail -c 300 gpurun_out/r04o_bench_ref.json
  python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r04o_smoke.log 2>&1; tail -2 gpurun_out/r04o_smoke.log
else
  ( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 ) > gpurun_out/r04o_bench_n$N.json 2> gpurun_out/r04o_bench_n$N.err
  tail -c 300 gpurun_out/r04o_bench_n$N.json; tail -3 gpurun_out/r04o_bench_n$N.err
fi
