#!/bin/bash
# round 2, call 55: part B as 4 x 256-thread or 2 x 512-thread blocks per SM (same 32 warps; fewer blocks to launch when nothing is listed)
out=gpurun_out; mkdir -p $out
bash scripts/r2_ab.sh r2c55 build/variants/lib_tpb256.so build/variants/lib_tpb512.so
for lib in visma_b200/libvisma_b200.so build/variants/lib_tpb256.so build/variants/lib_tpb512.so; do echo $lib; VISMA_B200_LIB=$PWD/$lib W=8 timeout 200 python scripts/time_shard_traj.py 2>&1 | tail -2; done | tee $out/r2c55_shard.txt
