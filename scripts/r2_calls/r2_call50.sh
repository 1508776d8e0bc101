#!/bin/bash
# round 2, call 50: where a vb200_icp_run-equivalent call spends its host time (scripts/time_e2e_phases.py, time_e2e.py)
out=gpurun_out; mkdir -p $out
timeout 300 python scripts/time_e2e_phases.py 2>&1 | tail -4 | tee $out/r2c50_phases.txt

