#!/bin/bash
# round 2, call 61: staged pageable copies: workers x chunk size
out=gpurun_out; mkdir -p $out
for lib in visma_b200/libvisma_b200.so build/variants/lib_w8c2.so build/variants/lib_w8c4.so build/variants/lib_w6c2.so build/variants/lib_w4c1.so; do echo "== $lib"; VISMA_B200_LIB=$PWD/$lib timeout 300 python scripts/time_scene_create.py 2>&1 | tail -3; done | tee $out/r2c61_scene.txt
nproc
