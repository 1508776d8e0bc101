#!/bin/bash
# round 2, call 64: cached proof of emptiness (an unmatched point keeps a limit on EVERY scene point from a search that looked
# half a fine cell beyond the radius, and skips its searches while it has moved less): GPU tests, config-2 per-object flow and
# bench line against the build before it
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $out/r2c64_pytest.log
for lib in visma_b200/libvisma_b200.so build/variants/lib_before.so; do echo "== $lib"; VISMA_B200_LIB=$PWD/$lib REPS=4 timeout 200 python scripts/profile_config2_traj.py 2>&1 | tail -3 | cut -c1-160; done | tee $out/r2c64_cfg2.txt
bash scripts/r2_ab.sh r2c64 build/variants/lib_before.so
