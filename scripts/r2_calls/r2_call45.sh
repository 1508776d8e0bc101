#!/bin/bash
# round 2, call 45: part B at 8 / 7 / 5 resident blocks per SM (64 / 72 / 96 registers) on top of the one-sided fences;
# ncu --set full of part B, iteration 1, pipelined-scan build (why it is not faster)
out=gpurun_out; mkdir -p $out
bash scripts/r2_ab.sh r2c45 build/variants/lib_lf.so build/variants/lib_lf_b8.so build/variants/lib_lf_b7.so build/variants/lib_lf_b5.so
VISMA_B200_LIB=$PWD/build/variants/lib_lf_pipe2.so N_ITER=2 timeout 300 ncu --set full --import-source on --clock-control none -k regex:"k_pass_b_wl" -s 1 -c 1 -o $out/r2c45_iter1_pipe2 python scripts/profile_traj.py > /dev/null 2>&1
ls -la $out/r2c45_iter1_pipe2.ncu-rep
