#!/bin/bash
# round 2, call 46: branch-free pipelined scan (run switches by selects, one instantiation) at 6 / 7 / 8 blocks per SM
out=gpurun_out; mkdir -p $out
bash scripts/r2_ab.sh r2c46 build/variants/lib_lf_b8.so build/variants/lib_lf_pipe3.so build/variants/lib_lf_pipe3_b7.so build/variants/lib_lf_pipe3_b8.so
N_ITER=8 timeout 200 python scripts/dump_trajectory.py /tmp/a.npz > /dev/null 2>&1
VISMA_B200_LIB=$PWD/build/variants/lib_lf_pipe3_b8.so N_ITER=8 timeout 200 python scripts/dump_trajectory.py /tmp/b.npz > /dev/null 2>&1
python scripts/dump_trajectory.py --cmp /tmp/a.npz /tmp/b.npz 2>&1 | tail -5 | tee $out/r2c46_cmp.txt
