#!/bin/bash
# round 2, call 62: launch list of ONE config-2 RegisterModelToScene (24 yaw starts): which kernels its 14 ms are
out=gpurun_out; mkdir -p $out
REPS=3 timeout 200 python scripts/profile_config2_traj.py 2>&1 | tail -3
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
timeout 300 ncu --metrics $M --clock-control none -k regex:"k_pass_a|k_pass_b_wl|k_solve" --csv --log-file $out/r2c62_cfg2.csv python scripts/profile_config2_traj.py > /dev/null 2>&1
python scripts/launch_table.py $out/r2c62_cfg2.csv | tee $out/r2c62_cfg2.txt | tail -45
