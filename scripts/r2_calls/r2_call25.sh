#!/bin/bash
# round 2, call 25: 64-byte hot records (match coordinates + normal streamed with the point): tests, A/B at 40/48/64/80 registers, 4-object launch list
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_icp.py tests/test_gpu_shard.py tests/test_gpu_dropin.py tests/test_gpu_config2.py tests/test_zz_gpu_real_pair.py -m gpu -x -q 2>&1 | tail -3
bash scripts/r2_ab.sh r2c25 build/variants/lib_b10.so build/variants/lib_b8.so build/variants/lib_b6.so build/variants/lib_ctl.so
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
N_ITER=30 STRIDE=8 timeout 300 ncu --metrics $M --clock-control none -k regex:"k_pass_a|k_pass_b_wl|k_solve" --csv --log-file $out/r2c25_launches_4obj.csv python scripts/profile_traj.py > $out/r2c25_b.log 2>&1
python scripts/launch_table.py $out/r2c25_launches_4obj.csv | tail -4
N_ITER=30 STRIDE=1 timeout 300 ncu --metrics $M --clock-control none -k regex:"k_pass_a|k_pass_b_wl|k_solve" --csv --log-file $out/r2c25_launches_32obj.csv python scripts/profile_traj.py > $out/r2c25_c.log 2>&1
python scripts/launch_table.py $out/r2c25_launches_32obj.csv | sed -n '1,8p;28,33p'
