#!/bin/bash
# round 2, call 48: second-line prefetch of long runs, no per-run drop thresholds, 12 runs per lane; one rank's share at
# 8 and 4 GPUs timed on one GPU (scripts/time_shard_traj.py) for the shipped build
out=gpurun_out; mkdir -p $out
bash scripts/r2_ab.sh r2c48 build/variants/lib_pf2.so build/variants/lib_nodrop.so build/variants/lib_runs12.so
for w in 8 4; do W=$w timeout 200 python scripts/time_shard_traj.py 2>&1 | tail -2; done | tee $out/r2c48_shard.txt
