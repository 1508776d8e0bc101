#!/bin/bash
# round 2, call 23: warp-cooperative LDL^T in k_solve, part B shortcuts (empty list, single-batch blocks)
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_icp.py tests/test_gpu_shard.py tests/test_gpu_dropin.py tests/test_gpu_config2.py tests/test_tools.py tests/test_zz_gpu_real_pair.py -m gpu -x -q 2>&1 | tail -3
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
N_ITER=30 STRIDE=8 timeout 300 ncu --metrics $M --clock-control none -k regex:"k_pass_a|k_pass_b_wl|k_solve" --csv --log-file $out/r2_launches_4obj_warp_solve.csv python scripts/profile_traj.py > $out/r2c23_b.log 2>&1
python scripts/launch_table.py $out/r2_launches_4obj_warp_solve.csv | tail -8
bash scripts/r2_ab.sh r2c23
