#!/bin/bash
# round 2, call 1: A/B of the round-1 pending variants + ncu captures of the SHIPPED pass kernels
out=gpurun_out; mkdir -p $out
TEST="" CMP="pretest nbw2 xyzn" bash scripts/ab_gpu_run.sh r2c1 pretest nbw2 nbw4 xyzn infl32 nonb
date -u +%T
# k_pass_b_wl: launches 0 and 1 after set_problems = iteration 0 (no prior) and iteration 1
N_ITER=3 timeout 200 ncu --set full --import-source on --clock-control none -k regex:"k_pass_b_wl" -s 0 -c 2 -o $out/r2_kpass_b_wl_iter01 python scripts/dump_trajectory.py /tmp/t.npz > $out/r2_ncu1.log 2>&1
tail -2 $out/r2_ncu1.log
# settled pass: launches 24.. of the plane trajectory
N_ITER=26 timeout 200 ncu --set full --import-source on --clock-control none -k regex:"k_pass_a|k_pass_b_wl|k_solve" -s 72 -c 3 -o $out/r2_settled python scripts/dump_trajectory.py /tmp/t.npz > $out/r2_ncu2.log 2>&1
tail -2 $out/r2_ncu2.log
date -u +%T
ls -la $out
