#!/bin/bash
# round 2, call 13: per-launch durations of the settled pass (new layout vs control) + one full capture of settled k_pass_a
out=gpurun_out; mkdir -p $out
for tag in new ctl; do
  lib=$PWD/visma_b200/libvisma_b200.so; [ $tag = ctl ] && lib=$PWD/build/variants/lib_ctl.so
  VISMA_B200_LIB=$lib N_ITER=26 timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_pass_a|k_pass_b_wl|k_solve" -s 60 -c 18 --csv --log-file $out/r2c13_launches_$tag.csv python scripts/dump_trajectory.py /tmp/t.npz > $out/r2c13_$tag.log 2>&1
done
N_ITER=26 timeout 300 ncu --set full --import-source on --clock-control none -k regex:"k_pass_a" -s 24 -c 1 -o $out/r2c13_kpass_a_settled python scripts/dump_trajectory.py /tmp/t.npz > $out/r2c13_ncu.log 2>&1
tail -2 $out/r2c13_ncu.log
