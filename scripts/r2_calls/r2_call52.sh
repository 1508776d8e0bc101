#!/bin/bash
# round 2, call 52: vb200_icp_run's two halves with programmatic dependent launch on both streams; no split at all
out=gpurun_out; mkdir -p $out
bash scripts/r2_ab.sh r2c52 build/variants/lib_e2e_chain.so build/variants/lib_e2e_nosplit.so
bash scripts/r2_ab.sh r2c52b build/variants/lib_e2e_chain.so build/variants/lib_e2e_nosplit.so
