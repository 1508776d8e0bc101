"""bench.py's JSON contract on the one arm that runs without a GPU: `--impl reference` (the unmodified
reference's CPU path from oracle/_ref), plus the roofline arithmetic.  CPU only; ~10 s."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT


def test_algorithmic_bytes_matches_survey():
    sys.path.insert(0, ROOT)
    import bench
    # N = 2 M, 32 x 50 k, all matched -> 32.0 + 38.4 + 89.6 MB + 216 B x 32
    b = bench.algorithmic_bytes(2_000_000, 32 * 50_000, 32 * 50_000, 32)
    assert b == 32_000_000 + 38_400_000 + 89_600_000 + 216 * 32
    assert bench.METRIC.startswith("icp_iterations_per_s") and bench.UNIT == "iterations/s"


def test_timed_steps_are_a_stratified_sample_of_the_trajectory():
    """`value` must not depend on --steps: K < 30 timed steps are spread evenly over the 30-iteration trajectory,
    K >= 30 times whole trajectories (the 20-step and the 30-step figures differed by 20 %)."""
    sys.path.insert(0, ROOT)
    import bench
    for k in (1, 3, 7, 20, 29, 30, 31, 45, 60, 100):
        plan = bench.trajectory_plan(k)
        assert sum(len(p) for p in plan) == k
        for p in plan:
            assert len(set(p)) == len(p) and all(0 <= x < 30 for x in p) and p == sorted(p)
    p20 = bench.trajectory_plan(20)[0]
    assert p20[0] == 0 and p20[-1] >= 28 and max(b - a for a, b in zip(p20, p20[1:])) <= 2
    assert bench.trajectory_plan(60) == [list(range(30))] * 2


def test_reference_arm_line(ref):
    env = dict(os.environ, RANK="0", WORLD_SIZE="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
              "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert k in line, k
    assert line["impl"] == "reference" and line["unit"] == "iterations/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "reference" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["value"] == line["value"] and line["e2e"]["h2d_bytes_per_step"] == 0
    assert "workload" in line["config"] and line["vs_baseline"] is None
    assert line["scaling"] == "strong" and line["config"]["objects"] == 32
    # ranks other than 0 do no work and print nothing
    env["RANK"] = "1"
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=120, env=env, cwd=ROOT)
    assert out.returncode == 0 and out.stdout.strip() == ""
