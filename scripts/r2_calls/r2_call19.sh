#!/bin/bash
# round 2, call 19: exhaustive KNN, compares chained into one predicate
out=gpurun_out; mkdir -p $out
timeout 240 python -m pytest tests/test_gpu_knn.py -m gpu -x -q 2>&1 | tail -3
timeout 120 python scripts/time_bruteforce.py 1e5 1e6 1e7 2>&1 | tee $out/r2c19_bruteforce.txt
timeout 300 ncu --set full --import-source on --clock-control none -k regex:"k_bf_knn1_f32" -s 1 -c 1 -o $out/r2c19_bf_f32 python scripts/time_bruteforce.py 1e6 > $out/r2c19_ncu.log 2>&1
