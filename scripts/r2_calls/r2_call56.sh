#!/bin/bash
# round 2, call 56: part B as 2 x 512 or 1 x 1024 threads per SM against 8 x 128, two runs each; one rank's share at 8 GPUs
out=gpurun_out; mkdir -p $out
bash scripts/r2_ab.sh r2c56 build/variants/lib_tpb512.so build/variants/lib_tpb1024.so
bash scripts/r2_ab.sh r2c56b build/variants/lib_tpb512.so build/variants/lib_tpb1024.so
for lib in visma_b200/libvisma_b200.so build/variants/lib_tpb512.so build/variants/lib_tpb1024.so visma_b200/libvisma_b200.so build/variants/lib_tpb512.so build/variants/lib_tpb1024.so; do echo $lib; VISMA_B200_LIB=$PWD/$lib W=8 timeout 200 python scripts/time_shard_traj.py 2>&1 | tail -2 | head -1; done | tee $out/r2c56_shard.txt
