#!/bin/bash
# The round's last short GPU call: the C++ drop-in harness and the C++ render_depth tool on the box, then
# compute-sanitizer (memcheck / racecheck / initcheck) over the kernels of the last session.
tag=${1:-r1s4b}
out=gpurun_out
mkdir -p $out
date -u +%T
timeout 120 python -m pytest tests/test_tools.py tests/test_gpu_dropin.py -m gpu -q > $out/${tag}_pytest_tools_dropin.log 2>&1; tail -4 $out/${tag}_pytest_tools_dropin.log
date -u +%T
for tool in memcheck racecheck initcheck; do
  echo "== $tool"; timeout 75 compute-sanitizer --tool $tool python scripts/sanitize_new_kernels.py 2>&1 | grep -v "^$" | tail -14
  date -u +%T
done > $out/${tag}_compute_sanitizer.txt 2>&1
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|^icp|^knn|^render" $out/${tag}_compute_sanitizer.txt
date -u +%T
