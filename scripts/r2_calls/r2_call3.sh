#!/bin/bash
out=gpurun_out; mkdir -p $out
date -u +%T
timeout 300 python -m pytest tests/test_gpu_icp.py tests/test_gpu_knn.py -m gpu -x -q > $out/r2c3_pytest.log 2>&1; tail -8 $out/r2c3_pytest.log
date -u +%T
VISMA_B200_LIB=$PWD/build/variants/lib_stats.so timeout 200 python scripts/search_stats.py 9 2>&1 | grep -v "^$" | cut -c1-600 > $out/r2c3_stats.txt; cat $out/r2c3_stats.txt
timeout 200 python bench.py --steps 30 --warmup 3 --no-cpu-baseline > $out/r2c3_bench.json 2> $out/r2c3_bench.err; python - <<'P'
import json
j=json.load(open("gpurun_out/r2c3_bench.json"))
c=j["config"]
print("value", j["value"], "pass_ms", c["pass_ms"], "solve", c["solve_ms"], "b2b", c["back_to_back_ms_per_iteration_no_flush"])
print(c["pass_ms_per_step"])
print("e2e", j["e2e"]["note"][:120]); print("default", j["e2e_default_criteria"]); print("abl", c["ablation_search_every_point_every_pass"])
P
tail -3 $out/r2c3_bench.err
date -u +%T
