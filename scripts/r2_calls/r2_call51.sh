#!/bin/bash
# round 2, call 51: vb200_icp_run with a live convergence test: both halves' four-iteration chunks enqueued before either
# finished-problems counter is read (the halves' tails overlap) against the sequential polls; ICP tests
out=gpurun_out; mkdir -p $out
bash scripts/r2_ab.sh r2c51 build/variants/lib_seqpoll.so
bash scripts/r2_ab.sh r2c51b build/variants/lib_seqpoll.so
timeout 900 python -m pytest tests/test_gpu_icp.py tests/test_gpu_config2.py tests/test_gpu_dropin.py -m gpu -x -q 2>&1 | tail -4 | tee $out/r2c51_pytest.log
