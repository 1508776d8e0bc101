#!/bin/bash
# round 2, call 42: software-pipelined candidate scan (next step's loads requested before this step is evaluated), the
# hand-out counter read one item ahead, no per-run L1 prefetch, lane-private reach 150 %, shared-walk threshold 16 lanes;
# bit-for-bit trajectory comparison of the pipelined scan against the shipped build
out=gpurun_out; mkdir -p $out
bash scripts/r2_ab.sh r2c42 build/variants/lib_pipe.so build/variants/lib_pipe_ip.so build/variants/lib_pipe_np.so build/variants/lib_pipe_r150.so build/variants/lib_pipe_c16.so
N_ITER=8 timeout 200 python scripts/dump_trajectory.py /tmp/a.npz > /dev/null 2>&1
VISMA_B200_LIB=$PWD/build/variants/lib_pipe_ip.so N_ITER=8 timeout 200 python scripts/dump_trajectory.py /tmp/b.npz > /dev/null 2>&1
python scripts/dump_trajectory.py --cmp /tmp/a.npz /tmp/b.npz 2>&1 | tail -5 | tee $out/r2c42_cmp.txt
