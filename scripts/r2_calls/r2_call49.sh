#!/bin/bash
# round 2, call 49: shared walk with staged candidates (one coalesced load per cell, shared-memory broadcasts) against the
# broadcast global loads; bit-for-bit trajectory comparison; KNN + ICP GPU tests on the staged build
out=gpurun_out; mkdir -p $out
bash scripts/r2_ab.sh r2c49 build/variants/lib_nostage.so
VISMA_B200_LIB=$PWD/build/variants/lib_nostage.so N_ITER=8 timeout 200 python scripts/dump_trajectory.py /tmp/a.npz > /dev/null 2>&1
N_ITER=8 timeout 200 python scripts/dump_trajectory.py /tmp/b.npz > /dev/null 2>&1
python scripts/dump_trajectory.py --cmp /tmp/a.npz /tmp/b.npz 2>&1 | tail -5 | tee $out/r2c49_cmp.txt
timeout 900 python -m pytest tests/test_gpu_icp.py tests/test_gpu_knn.py tests/test_gpu_config2.py -m gpu -x -q 2>&1 | tail -4 | tee $out/r2c49_pytest.log
