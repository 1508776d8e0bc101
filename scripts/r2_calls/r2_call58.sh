#!/bin/bash
# round 2, call 58: 36 / 40 part-B warps per SM as two blocks of 576 / 640 threads against 32 as one block of 1024 (two runs)
out=gpurun_out; mkdir -p $out
bash scripts/r2_ab.sh r2c58 build/variants/lib_w36b.so build/variants/lib_w40b.so
bash scripts/r2_ab.sh r2c58b build/variants/lib_w36b.so
for lib in visma_b200/libvisma_b200.so build/variants/lib_w36b.so; do echo $lib; VISMA_B200_LIB=$PWD/$lib W=8 timeout 200 python scripts/time_shard_traj.py 2>&1 | tail -2 | head -1; done | tee $out/r2c58_shard.txt
