#!/bin/bash
# round 2, call 16: exhaustive KNN with thresholds shared across threads and blocks: tests, timings, ncu
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_knn.py -m gpu -x -q 2>&1 | tail -5
timeout 300 python scripts/time_bruteforce.py 1e5 3e5 1e6 3e6 1e7 2>&1 | tee $out/r2c16_bruteforce.txt
timeout 300 ncu --set full --import-source on --clock-control none -k regex:"k_bf_knn1_f32" -s 1 -c 1 -o $out/r2c16_bf_f32 python scripts/time_bruteforce.py 1e6 > $out/r2c16_ncu.log 2>&1
tail -2 $out/r2c16_ncu.log
