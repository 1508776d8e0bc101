#!/bin/bash
# round 2, call 21: ncu evidence for the SHIPPED kernels: launch lists (32 objects and one rank's 4 at N = 8) and full captures
out=gpurun_out; mkdir -p $out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
N_ITER=30 STRIDE=1 timeout 300 ncu --metrics $M --clock-control none -k regex:"k_pass_a|k_pass_b_wl|k_solve" --csv --log-file $out/r2_launches_32obj.csv python scripts/profile_traj.py > $out/r2c21_a.log 2>&1
N_ITER=30 STRIDE=8 timeout 300 ncu --metrics $M --clock-control none -k regex:"k_pass_a|k_pass_b_wl|k_solve" --csv --log-file $out/r2_launches_4obj.csv python scripts/profile_traj.py > $out/r2c21_b.log 2>&1
# full captures: iteration 1 (launches 3, 4, 5 = pass A, pass B, solve) and iteration 24 (launches 72, 73, 74)
N_ITER=3 timeout 300 ncu --set full --import-source on --clock-control none -k regex:"k_pass_a|k_pass_b_wl|k_solve" -s 3 -c 3 -o $out/r2_iter01_shipped python scripts/profile_traj.py > $out/r2c21_c.log 2>&1
N_ITER=26 timeout 300 ncu --set full --import-source on --clock-control none -k regex:"k_pass_a|k_pass_b_wl|k_solve" -s 72 -c 3 -o $out/r2_iter24_shipped python scripts/profile_traj.py > $out/r2c21_d.log 2>&1
tail -1 $out/r2c21_a.log $out/r2c21_b.log $out/r2c21_c.log $out/r2c21_d.log
