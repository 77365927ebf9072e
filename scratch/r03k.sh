#!/bin/bash
# K1 per rank of 8 with the objects dealt out by distance from the camera (one GPU emulating each rank's shard)
mkdir -p gpurun_out
for r in 0 1 2 3 4 5 6 7; do
  ( timeout 120 python tools/sweep.py --workload c2 --frames 8 --what k1 --shard $r,8 --configs '[{}]' ) 2> /dev/null | grep '^{' | sed "s/^/{\"rank\": $r, \"line\": /; s/$/}/"
done > gpurun_out/r03k_k1_ranks_of8.jsonl
for r in 0 1 2 3; do
  ( timeout 120 python tools/sweep.py --workload c2 --frames 8 --what k1 --shard $r,4 --configs '[{}]' ) 2> /dev/null | grep '^{' | sed "s/^/{\"rank\": $r, \"line\": /; s/$/}/"
done > gpurun_out/r03k_k1_ranks_of4.jsonl
