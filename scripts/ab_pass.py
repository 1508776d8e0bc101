"""A/B the k_pass kernel time of alternative builds (dev tool): python scripts/ab_pass.py lib1.so lib2.so ..."""
import os, subprocess, sys, json
for lib in sys.argv[1:]:
    env = dict(os.environ, VISMA_B200_LIB=os.path.abspath(lib))
    out = subprocess.run([sys.executable, "bench.py", "--steps", "20", "--warmup", "3", "--no-cpu-baseline"],
                         env=env, capture_output=True, text=True)
    try:
        j = json.loads(out.stdout.strip().splitlines()[-1])
        print(lib, os.environ.get("VB200_CELL_SCALE", ""), "pass_ms %.4f" % j["config"]["pass_ms"], j["config"].get("pass_ms_first3"), j["config"].get("pass_ms_last3"), "value %.1f" % j["value"], "e2e %.1f" % j["e2e"]["value"], j["e2e"]["note"], j["config"].get("pass_ms_per_step"))
    except Exception as e:
        print(lib, "FAILED", out.stdout[-500:], out.stderr[-1500:])
