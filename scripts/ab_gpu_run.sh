#!/bin/bash
# dev tool: A/B alternative builds (build/variants/lib_<name>.so) on the GPU box in one short call:
#   bench line per build, the GPU tests ($TESTFILES, default ICP + KNN) on the builds named in $TEST, trajectory
#   comparison base vs each build named in $CMP.  Example (the variants prepared at the end of round 1):
#     scripts/build_variants.sh pretest "-DVB_PA_PRETEST" infl32 "-DVB_SOLVE_INFLIGHT=32" infl64 "-DVB_SOLVE_INFLIGHT=64" \
#         nbw2 "-DVB_NB_WINDOW=2" nbw4 "-DVB_NB_WINDOW=4" xyzn "-DVB_XYZN" xyzn_nohi8 "-DVB_XYZN -DVB_PA_NOHI -DVB_PASS_A_MINBLOCKS=8" perblock "-DVB_PB_PER_BLOCK"
#     gpurun --timeout 300 -- 'TEST="xyzn pretest" CMP="xyzn pretest infl32" bash scripts/ab_gpu_run.sh r2ab \
#         pretest infl32 infl64 nbw2 nbw4 xyzn xyzn_nohi8 perblock'
#   (build/variants/ travels with the snapshot: ~4 MB per build; delete it before other calls.)
tag=${1:-ab}; shift
out=gpurun_out; mkdir -p $out
libs="visma_b200/libvisma_b200.so"
for n in "$@"; do libs="$libs build/variants/lib_$n.so"; done
date -u +%T
timeout 150 python scripts/ab_pass.py $libs > $out/${tag}_ab.txt 2>&1; cut -c1-260 $out/${tag}_ab.txt
date -u +%T
for n in $TEST; do
  VISMA_B200_LIB=$PWD/build/variants/lib_$n.so timeout 90 python -m pytest ${TESTFILES:-tests/test_gpu_icp.py tests/test_gpu_knn.py} -m gpu -x -q 2>&1 | tail -2 | sed "s/^/$n: /"
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
