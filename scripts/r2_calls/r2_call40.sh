#!/bin/bash
# round 2, call 40: grid resolution with the row lister: occupancy target 3 / 5 / 9 points per occupied fine cell
out=gpurun_out; mkdir -p $out
bash scripts/r2_ab.sh r2c40 build/variants/lib_occ3.so build/variants/lib_occ9.so
