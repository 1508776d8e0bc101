#!/bin/bash
# round 2, call 29: with the cheaper row lister, re-tune when a warp takes the shared walk (coop lanes 12 / 20 / 32, big reach 75 / 100 %, lane reach 125 / 145 %)
out=gpurun_out; mkdir -p $out
bash scripts/r2_ab.sh r2c29 build/variants/lib_coop32.so build/variants/lib_coop20.so build/variants/lib_big100.so build/variants/lib_rho145.so
