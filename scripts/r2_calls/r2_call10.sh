#!/bin/bash
# round 2, call 10: part A with batched gathers / exact cached test, LDL^T with det from D: tests, then A/B vs control
out=gpurun_out; mkdir -p $out
date -u +%T
timeout 600 python -m pytest tests/test_gpu_icp.py tests/test_gpu_knn.py tests/test_gpu_dropin.py tests/test_gpu_shard.py -m gpu -x -q -k "not ten_million" 2>&1 | tail -5
date -u +%T
bash scripts/r2_ab.sh r2c10 build/variants/lib_g4b3.so build/variants/lib_g2b5.so build/variants/lib_ctl.so
date -u +%T
