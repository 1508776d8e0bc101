#!/bin/bash
# round 2, call 59: part A as 256-thread blocks x 2 batches per warp, or 64-thread blocks x 8, against 128 x 4 (same 512-point chunk)
out=gpurun_out; mkdir -p $out
bash scripts/r2_ab.sh r2c59 build/variants/lib_a256.so build/variants/lib_a64.so
for lib in visma_b200/libvisma_b200.so build/variants/lib_a256.so build/variants/lib_a64.so; do echo $lib; VISMA_B200_LIB=$PWD/$lib W=8 timeout 200 python scripts/time_shard_traj.py 2>&1 | tail -2 | head -1; done | tee $out/r2c59_shard.txt
