#!/bin/bash
# round 2, call 36: bench.py with the config-2 leg
out=gpurun_out; mkdir -p $out
date -u +%T
timeout 900 python bench.py --no-cpu-baseline > $out/r2c36_bench.json 2> $out/r2c36_bench.err
date -u +%T
python - <<P
import json
j=[json.loads(l) for l in open("gpurun_out/r2c36_bench.json") if l.startswith("{")][-1]
print("value", j["value"]); print(json.dumps(j.get("config2"), indent=1)[:2500])
P
tail -3 $out/r2c36_bench.err
