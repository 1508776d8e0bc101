#!/bin/bash
# round 2, call 2: correctness of the hot/cold-record pass + cluster solve, then a first bench and a sanitizer pass
out=gpurun_out; mkdir -p $out
date -u +%T
timeout 400 python -m pytest tests/test_gpu_icp.py tests/test_gpu_knn.py tests/test_gpu_dropin.py tests/test_annotation.py tests/test_tools.py -m gpu -x -q > $out/r2c2_pytest.log 2>&1; tail -15 $out/r2c2_pytest.log
date -u +%T
timeout 200 python bench.py --steps 30 --warmup 3 --no-cpu-baseline > $out/r2c2_bench.json 2> $out/r2c2_bench.err; tail -c 3000 $out/r2c2_bench.json; tail -5 $out/r2c2_bench.err
date -u +%T
echo "== memcheck"; timeout 120 compute-sanitizer --tool memcheck python scripts/sanitize_new_kernels.py 2>&1 | grep -v "^$" | tail -8
echo "== racecheck"; timeout 120 compute-sanitizer --tool racecheck python scripts/sanitize_new_kernels.py 2>&1 | grep -v "^$" | tail -6
date -u +%T
