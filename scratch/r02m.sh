#!/bin/bash
mkdir -p gpurun_out
CFG='[{"TGB_K1_DEFER_WORD":0,"TGB_K1_MIN_CTAS":4},{"TGB_K1_DEFER_WORD":0},{"TGB_K1_DEFER_WORD":1},{"TGB_K1_DEFER_WORD":1,"TGB_K1_MIN_CTAS":4},{"TGB_K1_DEFER_WORD":1,"TGB_K1_MIN_CTAS":6}]'
( time timeout 600 python tools/sweep.py --workload c2 --frames 10 --configs "$CFG" ) > gpurun_out/r02m_sweep_k1.jsonl 2> gpurun_out/r02m_sweep_k1.err; tail -3 gpurun_out/r02m_sweep_k1.err
( time timeout 600 python tools/sweep.py --workload c2far --frames 10 --configs "$CFG" ) > gpurun_out/r02m_sweep_k1_far.jsonl 2> gpurun_out/r02m_sweep_k1_far.err
