#!/bin/bash
# small knob checks with spare GPU time: pool CTAs per SM 9, K1 CTAs, k_shade is untouched
mkdir -p gpurun_out
CFG='[{},{"TGB_GI_POOL_CTAS_PER_SM":9},{"TGB_GI_POOL_CTAS_PER_SM":7},{"TGB_GI_POOL_CTAS_PER_SM":9,"TGB_GI_POOL_SERVICE_SLOTS":48},{"TGB_GI_POOL_SERVICE_SLOTS":48},{"TGB_GI_POOL_SERVICE_SLOTS":80},{"TGB_GI_POOL_DDA_STEPS":24},{"TGB_GI_POOL_DDA_STEPS":12},{"TGB_GI_POOL_TREE_REPS":3},{"TGB_GI_POOL_TREE_REPS":6}]'
( time timeout 600 python tools/sweep.py --workload c2 --frames 12 --configs "$CFG" ) > gpurun_out/r03m_sweep_pool.jsonl 2> gpurun_out/r03m_sweep_pool.err
tail -2 gpurun_out/r03m_sweep_pool.err
