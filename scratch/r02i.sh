#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02i_pytest_gpu.log 2>&1; echo "exit $?" >> gpurun_out/r02i_pytest_gpu.log; tail -12 gpurun_out/r02i_pytest_gpu.log
