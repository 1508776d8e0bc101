#!/bin/bash
# round 2, call 43: part B's per-batch __threadfence() (fence.sc = MEMBAR.SC + CCTL.IVALL: the SM's whole L1 invalidated once
# per batch) replaced by fence.release.gpu on the writer and fence.acquire.gpu on the one folding batch; alone and with the
# pipelined scan / item prefetch / no run prefetch; bit-for-bit trajectory comparison against the shipped build
out=gpurun_out; mkdir -p $out
bash scripts/r2_ab.sh r2c43 build/variants/lib_lf.so build/variants/lib_lf_pipe.so build/variants/lib_lf_pipe_ip.so build/variants/lib_lf_np.so
N_ITER=8 timeout 200 python scripts/dump_trajectory.py /tmp/a.npz > /dev/null 2>&1
VISMA_B200_LIB=$PWD/build/variants/lib_lf_pipe_ip.so N_ITER=8 timeout 200 python scripts/dump_trajectory.py /tmp/b.npz > /dev/null 2>&1
python scripts/dump_trajectory.py --cmp /tmp/a.npz /tmp/b.npz 2>&1 | tail -5 | tee $out/r2c43_cmp.txt
