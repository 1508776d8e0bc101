#!/bin/bash
# round 2, call 47: new default (one-sided fences, 32 part-B warps per SM) against 36 / 40 / 48 warps per SM; ICP / shard GPU tests
out=gpurun_out; mkdir -p $out
bash scripts/r2_ab.sh r2c47 build/variants/lib_w36.so build/variants/lib_w40.so build/variants/lib_w48.so
timeout 900 python -m pytest tests/test_gpu_icp.py tests/test_gpu_shard.py tests/test_zz_gpu_real_pair.py tests/test_gpu_config2.py -m gpu -x -q 2>&1 | tail -4 | tee $out/r2c47_pytest.log
