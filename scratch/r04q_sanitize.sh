#!/bin/bash
# compute-sanitizer over the small-scene GPU tests of the final tree (memcheck on everything, racecheck on the shading tests: shared-memory pools, cp.async staging)
mkdir -p gpurun_out
SEL="not config2 and not config3 and not config4 and not config5 and not full_size"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 99 --print-limit 20 python -m pytest tests/test_visibility_gpu.py tests/test_svo_gpu.py tests/test_shading_gpu.py tests/test_procedural_gpu.py -m gpu -x -q -k "$SEL" > gpurun_out/r04q_memcheck.log 2>&1; echo "exit $?" >> gpurun_out/r04q_memcheck.log
tail -6 gpurun_out/r04q_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 99 --print-limit 20 python -m pytest tests/test_shading_gpu.py -m gpu -x -q -k "$SEL and (fast_walk or full_frame or golden)" > gpurun_out/r04q_racecheck.log 2>&1; echo "exit $?" >> gpurun_out/r04q_racecheck.log
tail -6 gpurun_out/r04q_racecheck.log
