#!/bin/bash
# round 2, call 44: pipelined scan with two alternating register sets (no copies between steps) on top of the one-sided
# fences; ncu --set full of iterations 0 and 1 of the one-sided-fence build
out=gpurun_out; mkdir -p $out
bash scripts/r2_ab.sh r2c44 build/variants/lib_lf.so build/variants/lib_lf_pipe2.so
VISMA_B200_LIB=$PWD/build/variants/lib_lf.so N_ITER=2 timeout 300 ncu --set full --import-source on --clock-control none -k regex:"k_pass_a|k_pass_b_wl|k_solve" -c 6 -o $out/r2c44_iter0_1_lf python scripts/profile_traj.py > /dev/null 2>&1
ls -la $out/r2c44_iter0_1_lf.ncu-rep
