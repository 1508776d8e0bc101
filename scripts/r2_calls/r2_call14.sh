#!/bin/bash
# round 2, call 14: part B block size (128 / 384 / 768 threads) and part A chunk (512 / 256 points per block)
out=gpurun_out; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_icp.py -m gpu -x -q -k "not full_size" 2>&1 | tail -3
bash scripts/r2_ab.sh r2c14 build/variants/lib_pb768.so build/variants/lib_pb384.so build/variants/lib_pts2.so
