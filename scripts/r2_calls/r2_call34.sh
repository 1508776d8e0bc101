#!/bin/bash
# round 2, call 34: part A trimmed (no square root in the cached test, no second distance / record load, rows staged straight to
# shared memory, per-problem all-inside flag): tests, A/B vs HEAD, and at 48 registers
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_icp.py tests/test_gpu_shard.py tests/test_gpu_dropin.py -m gpu -x -q 2>&1 | tail -3
bash scripts/r2_ab.sh r2c34 build/variants/lib_b10.so build/variants/lib_ctl.so
