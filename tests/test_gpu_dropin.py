"""The C++ host adapters (visma_b200/host/*.h) compiled against the reference's real Open3D headers and run
next to the reference's own CPU functions (oracle/_ref/dropin_check, built by `make -C oracle dropin`).
Covers BASELINE config 1 (plumbing): the reference's CPU RegistrationICP loop driving the GPU estimator."""
import json
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, small_scene

pytestmark = pytest.mark.gpu

BIN = os.path.join(ROOT, "oracle", "_ref", "dropin_check")


@pytest.mark.skipif(not os.path.exists(BIN), reason="oracle/_ref/dropin_check not built (needs /root/reference)")
def test_cpp_dropin_against_reference(tmp_path):
    d = small_scene(n_scene=100000, n_objects=2, m=50000, seed=21)  # config 1: one chair, one 50k fragment
    src, sn = d["sources"][0]
    f = tmp_path / "in.bin"
    with open(f, "wb") as fh:
        np.array([len(d["scene_xyz"]), len(src)], np.int64).tofile(fh)
        for a in (d["scene_xyz"], d["scene_nrm"], src, sn, d["T_init"][0]):
            np.ascontiguousarray(a, np.float64).tofile(fh)
    out = subprocess.run([BIN, str(f)], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr
    line = [l for l in out.stdout.splitlines() if l.startswith("JSON:")][-1]
    r = json.loads(line[5:])
    for k in ("p2p", "p2plane"):
        assert r[k]["dT"] < 1e-6, r[k]
        assert abs(r[k]["ncorr_ref"] - r[k]["ncorr_gpu"]) <= 2 and abs(r[k]["rmse_ref"] - r[k]["rmse_gpu"]) < 1e-6
    assert r["cicp_plugin"]["dT"] < 1e-9 and r["cicp_plugin"]["ncorr_ref"] == r["cicp_plugin"]["ncorr_mix"]
    # the gravity estimator class: CPU loop + GPU estimator == GPU loop; no roll / pitch
    assert r["gravity"]["dT"] < 1e-6 and abs(r["gravity"]["ncorr_mix"] - r["gravity"]["ncorr_gpu"]) <= 2
    assert r["gravity"]["roll_pitch"] < 1e-12
    # A1: ComputeRMSE on the GPU vs the reference's (summation order differs)
    assert abs(r["rmse"]["ref"] - r["rmse"]["gpu"]) < 1e-12 * max(1.0, r["rmse"]["ref"]) and r["rmse"]["empty"] == 0
    assert r["rmse"]["ref"] > 0
    # per-source normals decision in a mixed point-to-plane batch
    m = r["mixed_normals"]
    assert m["with_dT"] == 0 and m["with2_dT"] == 0 and m["with_ncorr"] > 1000
    assert m["bare_dT"] == 0 and m["bare_fitness"] == 0 and m["empty_dT"] == 0
    assert r["errors"]["bad_distance_dT"] == 0 and r["errors"]["no_normals_dT"] == 0
    assert r["errors"]["no_normals_fitness"] == 0
    assert r["voxel"]["n_ref"] == r["voxel"]["n_gpu"] and r["voxel"]["dsum"] < 1e-6
    assert r["render"]["covered"] > 10000 and abs(r["render"]["z_lin"] - 1.0) < 1e-3


SHARDED = os.path.join(ROOT, "oracle", "_ref", "sharded_check")


@pytest.mark.skipif(not os.path.exists(SHARDED), reason="oracle/_ref/sharded_check not built (needs /root/reference)")
def test_cpp_sharded_host_path_equals_single_gpu(tmp_path):
    """visma_b200::RegistrationICPSharded (C++ host, ncclAllGather of the 160-byte pose rows) on the
    GPUs this box has — two ranks when there are two devices, else one — against the single-GPU batch of all the
    objects: identical transforms on every rank."""
    d = small_scene(n_scene=120000, n_objects=5, m=5000, seed=31)
    f = tmp_path / "in.bin"
    with open(f, "wb") as fh:
        np.array([len(d["scene_xyz"]), len(d["sources"]), 5000], np.int64).tofile(fh)
        np.ascontiguousarray(d["scene_xyz"], np.float64).tofile(fh)
        np.ascontiguousarray(d["scene_nrm"], np.float64).tofile(fh)
        for p, n in d["sources"]:
            np.ascontiguousarray(p, np.float64).tofile(fh)
            np.ascontiguousarray(n, np.float64).tofile(fh)
        np.ascontiguousarray(d["T_init"], np.float64).tofile(fh)
    out = subprocess.run([SHARDED, str(f)], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, (out.returncode, out.stderr[-2000:])
    r = json.loads([l for l in out.stdout.splitlines() if l.startswith("JSON:")][-1][5:])
    assert r["objects"] == 5 and r["world"] >= 1
    assert r["max_dT"] == 0 and r["max_dfit"] == 0 and r["ncorr_mismatches"] == 0
    assert r["own_correspondence_sets_equal"] == 5 and r["fitness0"] > 0.9
    # RegistrationICPGlobalSharded: one cloud cut into slices, per-iteration ncclAllReduce of the totals — the same
    # transform on every rank, equal to the single-GPU alignment up to the summation order of the totals
    assert r["global_rank_spread"] == 0 and r["global_fitness"] > 0.9
    assert r["global_max_dT"] < 1e-9 and r["global_max_dfit"] < 1e-9
