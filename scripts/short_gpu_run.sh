#!/bin/bash
# A short GPU-box call (a few minutes): GPU tests, the render / KNN-sweep benches of the newest kernels, the
# bench line, then ncu captures while time remains.  Every step is bounded; outputs: gpurun_out/<tag>_*.
tag=${1:-r1s4}
out=gpurun_out
mkdir -p $out
date -u +%T
timeout 330 python -m pytest tests -m gpu -q > $out/${tag}_pytest_gpu.log 2>&1; tail -4 $out/${tag}_pytest_gpu.log
date -u +%T
timeout 60 python scripts/bench_render.py > $out/${tag}_render_bench.json 2> $out/${tag}_render.err; tail -c 600 $out/${tag}_render_bench.json; echo
timeout 90 python scripts/bench_knn_sweep.py > $out/${tag}_knn_sweep.json 2> $out/${tag}_knn_sweep.err; tail -c 300 $out/${tag}_knn_sweep.json; echo
date -u +%T
timeout 240 python bench.py > $out/${tag}_bench_1gpu.json 2> $out/${tag}_bench_1gpu.err; tail -c 600 $out/${tag}_bench_1gpu.json; echo
date -u +%T
timeout 90 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 120 ncu --set full --import-source on --clock-control none -k regex:"k_tris" -s 1 -c 1 -o $out/${tag}_ktris python scripts/bench_render.py > /dev/null 2>&1
timeout 120 ncu --set full --import-source on --clock-control none -k regex:"k_bf_knn1|k_knn1_wpq" -c 4 -o $out/${tag}_knn python scripts/profile_bruteforce.py > /dev/null 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $out/${tag}_launches_ncu.csv python bench.py --steps 30 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
date -u +%T
ls -la $out | grep ${tag}
