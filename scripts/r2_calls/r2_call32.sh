#!/bin/bash
# round 2, call 32: the lane's runs sorted nearest first before the scan (drop thresholds prune the far ones) vs box order
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_icp.py tests/test_gpu_knn.py -m gpu -x -q -k "not ten_million" 2>&1 | tail -3
bash scripts/r2_ab.sh r2c32 build/variants/lib_nosort.so
