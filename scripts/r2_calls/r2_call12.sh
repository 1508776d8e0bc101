#!/bin/bash
# round 2, call 12: 32-byte scene records (256-bit loads), SoA source points, exact cached test; tests then A/B
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_icp.py tests/test_gpu_knn.py tests/test_gpu_dropin.py tests/test_gpu_shard.py tests/test_gpu_config2.py -m gpu -x -q -k "not ten_million" 2>&1 | tail -5
bash scripts/r2_ab.sh r2c12 build/variants/lib_b10.so build/variants/lib_b8.so build/variants/lib_ctl.so
