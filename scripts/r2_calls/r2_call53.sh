#!/bin/bash
# round 2, call 53: part A's per-lane running sums in shared memory instead of (spilled) local memory; 48-register build;
# bit-for-bit trajectory comparison
out=gpurun_out; mkdir -p $out
bash scripts/r2_ab.sh r2c53 build/variants/lib_pasm.so build/variants/lib_pa10.so
N_ITER=10 timeout 200 python scripts/dump_trajectory.py /tmp/a.npz > /dev/null 2>&1
VISMA_B200_LIB=$PWD/build/variants/lib_pasm.so N_ITER=10 timeout 200 python scripts/dump_trajectory.py /tmp/b.npz > /dev/null 2>&1
python scripts/dump_trajectory.py --cmp /tmp/a.npz /tmp/b.npz 2>&1 | tail -5 | tee $out/r2c53_cmp.txt
