#!/bin/bash
# round 2, call 41: slack of the settling searches: upper cap 20 / 35 % of a fine cell, settling threshold 20 / 35 %, slack = 60 / 100 / 150 % of the latest move
out=gpurun_out; mkdir -p $out
bash scripts/r2_ab.sh r2c41 build/variants/lib_hi35.so build/variants/lib_set35.so build/variants/lib_mv150.so build/variants/lib_mv60.so
