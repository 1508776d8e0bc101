#!/bin/bash
# round 2, call 38: every allocation from the stream-ordered pool (voxel, sampler, scene build / destroy, bbox, host KNN entry): all GPU tests,
# the per-object phases of the AnnotationTool flow, the bench line
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 300 python scripts/time_annotate_phases.py 2>&1 | tail -4
timeout 600 python bench.py --no-cpu-baseline > $out/r2c38_bench.json 2> $out/r2c38_bench.err
python - <<P
import json
j=[json.loads(l) for l in open("gpurun_out/r2c38_bench.json") if l.startswith("{")][-1]
print("value", j["value"], "e2e", j["e2e"]["value"], "with_build", j["e2e_with_scene_build"]["ms_per_call"], "config2 ms/object", j["config2"]["ms_per_object"], j["config2"].get("parity_vs_cpu_flow_object0"))
print([ (r["N"], round(r["scene_create_s_incl_h2d"],4)) for r in j["knn_sweep"]["rows"]])
P
