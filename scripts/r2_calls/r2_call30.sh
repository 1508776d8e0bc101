#!/bin/bash
# round 2, call 30: vb200_icp_run's first pipeline piece = 1/2, 1/3, 1/4, 1/8 of the points (what the exposed upload costs)
out=gpurun_out; mkdir -p $out
bash scripts/r2_ab.sh r2c30 build/variants/lib_div3.so build/variants/lib_div4.so build/variants/lib_div8.so
