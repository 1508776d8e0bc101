#!/bin/bash
# round 2, call 63: the empty-space field (no search for an unmatched point whose home cell is provably farther than the
# radius from every occupied cell) against the build without it: config-2 per-object flow, bench line, GPU tests
out=gpurun_out; mkdir -p $out
for lib in visma_b200/libvisma_b200.so build/variants/lib_nogap.so; do echo "== $lib"; VISMA_B200_LIB=$PWD/$lib REPS=4 timeout 200 python scripts/profile_config2_traj.py 2>&1 | tail -3 | cut -c1-160; done | tee $out/r2c63_cfg2.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee $out/r2c63_pytest.log
bash scripts/r2_ab.sh r2c63 build/variants/lib_nogap.so
