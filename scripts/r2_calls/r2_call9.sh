#!/bin/bash
out=gpurun_out; mkdir -p $out
date -u +%T
timeout 300 python -m pytest tests/test_gpu_icp.py tests/test_gpu_knn.py -m gpu -x -q -k "not full_size and not ten_million" 2>&1 | tail -3
date -u +%T
bash scripts/r2_ab.sh r2c9 build/variants/lib_ctl.so
date -u +%T
timeout 1200 python -m pytest tests -m gpu -q > $out/r2_pytest_gpu_full.log 2>&1; tail -6 $out/r2_pytest_gpu_full.log
date -u +%T
