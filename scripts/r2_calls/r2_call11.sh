#!/bin/bash
# round 2, call 11: part A = L2-prefetch-chained walk with the exact cached test; A/B vs no-prefetch / 48-register / control
out=gpurun_out; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_icp.py -m gpu -x -q -k "not full_size" 2>&1 | tail -3
bash scripts/r2_ab.sh r2c11 build/variants/lib_pfb10.so build/variants/lib_nopf.so build/variants/lib_ctl.so
