#!/bin/bash
# round 2, call 57: part B as one 1024-thread block per SM (adopted): bit-for-bit trajectory against the 8 x 128 build, GPU tests
out=gpurun_out; mkdir -p $out
VISMA_B200_LIB=$PWD/build/variants/lib_tpb128.so N_ITER=10 timeout 200 python scripts/dump_trajectory.py /tmp/a.npz > /dev/null 2>&1
N_ITER=10 timeout 200 python scripts/dump_trajectory.py /tmp/b.npz > /dev/null 2>&1
python scripts/dump_trajectory.py --cmp /tmp/a.npz /tmp/b.npz 2>&1 | tail -5 | tee $out/r2c57_cmp.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $out/r2c57_pytest.log
for w in 4 2; do W=$w timeout 200 python scripts/time_shard_traj.py 2>&1 | tail -2 | head -1; VISMA_B200_LIB=$PWD/build/variants/lib_tpb128.so W=$w timeout 200 python scripts/time_shard_traj.py 2>&1 | tail -2 | head -1; done | tee $out/r2c57_shard.txt
