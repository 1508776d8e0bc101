#!/bin/bash
out=gpurun_out; mkdir -p $out
date -u +%T
timeout 300 python -m pytest tests/test_gpu_shard.py tests/test_gpu_dropin.py -m gpu -x -q > $out/r2c7_pytest.log 2>&1; tail -5 $out/r2c7_pytest.log
date -u +%T
timeout 600 python bench.py --steps 20 --warmup 5 > $out/r2c7_bench.json 2> $out/r2c7_bench.err; tail -3 $out/r2c7_bench.err; python - <<'P'
import json
j=json.load(open("gpurun_out/r2c7_bench.json")); c=j["config"]
print("value %.0f ms/step %.4f e2e %.0f launches %d" % (j["value"], j["ms_per_step"], j["e2e"]["value"], j["gpu_launches"]))
print("timed", c["timed_steps"]); print("step_ms", c["step_ms_timed"]); print("traj", c["trajectory_ms_per_step"])
print("regimes", c["search_bound_regime"], c["settled_regime"]); print("roofline", {k: j["roofline"][k] for k in ("achieved","frac","traffic")}, j["roofline"]["search_bound"]["frac"], j["roofline"]["settled"]["frac"])
print("e2e", j["e2e"]["note"]); print(j["e2e_default_criteria"]); print(j["e2e_with_scene_build"]); print("cpu", j["cpu_baseline"])
for r in j.get("knn_sweep",{}).get("rows",[]): print({k:(round(v,4) if isinstance(v,float) else v) for k,v in r.items() if k not in ("algorithmic_bytes","radius","Q")})
print("render", j.get("render"))
print("clocks", j["clocks"])
P
date -u +%T
