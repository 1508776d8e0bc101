#!/bin/bash
# round 2, call 35: point + normal records interleaved (64 B per scene point: one half line per gather instead of two lines)
out=gpurun_out; mkdir -p $out
VISMA_B200_LIB=$PWD/build/variants/lib_xn.so timeout 600 python -m pytest tests/test_gpu_icp.py -m gpu -x -q -k "not full_size" 2>&1 | tail -2
bash scripts/r2_ab.sh r2c35 build/variants/lib_xn.so
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
VISMA_B200_LIB=$PWD/build/variants/lib_xn.so N_ITER=26 STRIDE=1 timeout 300 ncu --metrics $M --clock-control none -k regex:"k_pass_a" -s 20 -c 4 --csv --log-file $out/r2c35_xn.csv python scripts/profile_traj.py > /dev/null 2>&1
grep "k_pass_a" $out/r2c35_xn.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}' | head -12
