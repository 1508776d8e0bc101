#!/bin/bash
# round 2, call 27: ncu capture of part B, iteration 1, row lister
out=gpurun_out; mkdir -p $out
N_ITER=3 timeout 300 ncu --set full --import-source on --clock-control none -k regex:"k_pass_b_wl" -s 1 -c 1 -o $out/r2c27_pb_iter1 python scripts/profile_traj.py > $out/r2c27.log 2>&1
tail -n 2 $out/r2c27.log
