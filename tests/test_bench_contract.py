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
    # ranks other than 0 do no work and print nothing
    env["RANK"] = "1"
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=120, env=env, cwd=ROOT)
    assert out.returncode == 0 and out.stdout.strip() == ""
