#!/bin/bash
# dev tool: A/B alternative builds (build/variants/lib_<name>.so) on the GPU box in one short call:
#   bench line per build, the ICP GPU tests on the builds named in $TEST, trajectory comparison base vs $CMP.
tag=${1:-ab}; shift
out=gpurun_out; mkdir -p $out
libs="visma_b200/libvisma_b200.so"
for n in "$@"; do libs="$libs build/variants/lib_$n.so"; done
date -u +%T
timeout 150 python scripts/ab_pass.py $libs > $out/${tag}_ab.txt 2>&1; cut -c1-260 $out/${tag}_ab.txt
date -u +%T
for n in $TEST; do
  VISMA_B200_LIB=$PWD/build/variants/lib_$n.so timeout 60 python -m pytest tests/test_gpu_icp.py -m gpu -x -q 2>&1 | tail -2 | sed "s/^/$n: /"
done | tee $out/${tag}_tests.txt
date -u +%T
if [ -n "$CMP" ]; then
  N_ITER=8 timeout 60 python scripts/dump_trajectory.py /tmp/base.npz | tail -1
  for n in $CMP; do
    VISMA_B200_LIB=$PWD/build/variants/lib_$n.so N_ITER=8 timeout 60 python scripts/dump_trajectory.py /tmp/$n.npz | tail -1
    python scripts/dump_trajectory.py --cmp /tmp/base.npz /tmp/$n.npz 2>&1 | tail -4 | sed "s/^/$n vs base: /"
  done | tee $out/${tag}_cmp.txt
fi
date -u +%T
