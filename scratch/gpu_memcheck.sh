#!/bin/bash
mkdir -p gpurun_out
cd /root/repo
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 99 --print-limit 20 python -m pytest tests/test_visibility_gpu.py tests/test_svo_gpu.py tests/test_shading_gpu.py tests/test_procedural_gpu.py -m gpu -x -q -k "not config2 and not config3 and not config4 and not config5 and not full_size" > gpurun_out/sanitize.log 2>&1; echo "exit $?" >> gpurun_out/sanitize.log
tail -15 gpurun_out/sanitize.log
grep -c "Invalid\|out of bounds\|misaligned" gpurun_out/sanitize.log
