#!/bin/bash
# round 2, call 54: part A's block length fitted to a whole number of waves (480 instead of 512 points per block on the
# BASELINE workload: 1.89 instead of 1.77 waves) against the fixed 512; ICP + shard tests
out=gpurun_out; mkdir -p $out
bash scripts/r2_ab.sh r2c54 build/variants/lib_nofit.so
bash scripts/r2_ab.sh r2c54b build/variants/lib_nofit.so
timeout 900 python -m pytest tests/test_gpu_icp.py tests/test_gpu_shard.py tests/test_gpu_config2.py -m gpu -x -q 2>&1 | tail -4 | tee $out/r2c54_pytest.log
