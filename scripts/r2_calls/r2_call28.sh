#!/bin/bash
# round 2, call 28: row lister with cheaper square roots and conversions
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_icp.py tests/test_gpu_knn.py tests/test_gpu_shard.py -m gpu -x -q -k "not ten_million" 2>&1 | tail -3
bash scripts/r2_ab.sh r2c28 build/variants/lib_ctl.so
