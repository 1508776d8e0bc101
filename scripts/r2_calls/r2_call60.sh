#!/bin/bash
# round 2, call 60: pageable host-to-device copies staged by four host threads through pinned slots (vb::h2d_async) against
# cudaMemcpyAsync from pageable memory: vb200_scene_create at 0.3-5 M points, bench line (e2e_with_scene_build), GPU tests
out=gpurun_out; mkdir -p $out
echo "== staged"; timeout 300 python scripts/time_scene_create.py 2>&1 | tail -4 | tee $out/r2c60_scene.txt
echo "== plain cudaMemcpyAsync"; VISMA_B200_LIB=$PWD/build/variants/lib_plaincopy.so timeout 300 python scripts/time_scene_create.py 2>&1 | tail -4 | tee -a $out/r2c60_scene.txt
bash scripts/r2_ab.sh r2c60 build/variants/lib_plaincopy.so
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee $out/r2c60_pytest.log
