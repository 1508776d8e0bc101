"""Host math of the annotation driver mirror (src/annotation.cpp:78-96, include/geometry.h:18-26,
core/utils.h:229-233) — CPU only — and, on the GPU, the whole orientation-constrained flow and the
ICPRefinement flow against the same flows driven by the oracle."""
import os

import numpy as np
import pytest

from conftest import small_scene


def test_plane_normal_and_gravity_rotation():
    from visma_b200 import annotation as an, synth
    rng = np.random.default_rng(0)
    Rg = synth.rot_xyz(0.2, -0.4, 0.1)
    floor = np.stack([rng.uniform(0, 5, 4000), rng.normal(0, 0.002, 4000), rng.uniform(0, 5, 4000)], 1) @ Rg.T
    n = an.FindPlaneNormal(floor)
    assert np.allclose(n, Rg @ [0, 1, 0], atol=2e-3) or np.allclose(n, -(Rg @ [0, 1, 0]), atol=2e-3)
    T0 = an.GravityAlignment(floor)
    assert np.allclose(T0[:3, :3] @ n, [0, 1, 0], atol=1e-12)
    assert np.allclose(T0[:3, :3] @ T0[:3, :3].T, np.eye(3), atol=1e-12) and np.linalg.det(T0[:3, :3]) > 0.999
    for u, v in (([1, 0, 0], [0, 1, 0]), ([0, 1, 0], [0, 1, 0]), ([0, -1, 0], [0, 1, 0]), ([1, 2, 3], [-3, 1, 2])):
        R = an.RotationBetweenVectors(u, v)
        assert np.allclose(R @ (np.array(u, float) / np.linalg.norm(u)), np.array(v, float) / np.linalg.norm(v))
        assert np.allclose(R @ R.T, np.eye(3), atol=1e-12)
    assert an.MinY(floor) == floor[:, 1].min()


@pytest.mark.gpu
def test_annotation_flow_matches_oracle_flow(vb, oracle):
    """AnnotationTool end to end: tilted scene -> gravity alignment -> VoxelDownSample -> centring -> yaw
    search -> composed pose; the GPU flow must equal the identical flow with the oracle's
    RegisterModelToScene, and land on the ground truth."""
    from visma_b200 import annotation as an
    d = small_scene(n_scene=400000, n_objects=2, m=3000, seed=13)
    Rg = vb.synth.rot_xyz(0.15, 0.3, -0.1)            # the scan's frame is not gravity aligned
    tilt = lambda p: p @ Rg.T
    xyz = d["scene_xyz"]
    floor = tilt(xyz[(np.abs(xyz[:, 1]) < 0.01) & (xyz[:, 0] < 1.0) & (xyz[:, 2] < 1.0)])
    cfg = {"ICP": {"voxel_size": 0.01, "point_to_plane": False, "rotation_level": 8, "distance_threshold": 0.02}}
    scans, models = {}, {}
    for b in range(2):
        c = d["T_gt"][b][:3, 3]
        near = (np.linalg.norm(xyz[:, [0, 2]] - c[[0, 2]], axis=1) < 0.6) & (xyz[:, 1] > 0.02)
        scans["chair%d_0" % b] = tilt(xyz[near])
        models["chair%d" % b] = d["sources"][b][0]
    reg_oracle = lambda model, scan: oracle.register_model_to_scene(model, scan, level=8, threshold=0.02)
    got = an.AnnotationTool(floor, scans, models, cfg)
    ref = an.AnnotationTool(floor, scans, models, cfg, register=reg_oracle)
    for b, name in enumerate(sorted(got)):
        assert got[name].shape == (3, 4)
        assert np.allclose(got[name], ref[name], atol=1e-6), name
        Tgt = np.eye(4)
        Tgt[:3, :3] = Rg @ d["T_gt"][b][:3, :3]
        Tgt[:3, 3] = Rg @ d["T_gt"][b][:3, 3]
        T = np.eye(4)
        T[:3] = got[name]
        rot, tr = vb.synth.pose_error(T, Tgt)
        assert rot < 0.05 and tr < 0.03, (name, rot, tr)


@pytest.mark.gpu
def test_icp_refinement_flow(vb, oracle):
    """feh::ICPRefinement: union of posed model samples, voxel-down-sampled scene, one global ICP."""
    from visma_b200 import annotation as an
    d = small_scene(n_scene=300000, n_objects=4, m=5000, seed=17)
    clouds = [p for p, _ in d["sources"]]
    off = vb.synth.make_T(vb.synth.rot_xyz(0.004, -0.01, 0.003), [0.01, -0.004, 0.012])  # global misalignment
    opts = {"voxel_size": 0.02, "max_distance": 0.075, "use_point_to_plane": False}
    res, scene = an.ICPRefinement(d["scene_xyz"], clouds, list(d["T_gt"]), off, opts)
    est = np.concatenate([p @ T[:3, :3].T + T[:3, 3] for p, T in zip(clouds, d["T_gt"])])
    o_scene = oracle.voxel_downsample(d["scene_xyz"], 0.02)
    assert (scene.points_ == o_scene).all()
    o = oracle.Index(o_scene, 0.075).registration_icp(est, 0.075, off, oracle.P2P)
    rot, tr = vb.synth.pose_error(res.transformation_, o["T"])
    assert rot < 1e-6 and tr < 1e-6
    assert abs(res.fitness_ - o["fitness"]) <= 2.0 / len(est)
    rot, tr = vb.synth.pose_error(res.transformation_, np.eye(4))  # and it undid the misalignment
    assert rot < 5e-3 and tr < 5e-3
    # the reference's point-to-plane branch returns the init unchanged: scene_est has no normals
    opts["use_point_to_plane"] = True
    res, _ = an.ICPRefinement(d["scene_xyz"], clouds, list(d["T_gt"]), off, opts)
    assert np.array_equal(res.transformation_, off) and res.fitness_ == 0


def test_savemat_format(tmp_path):
    """feh::SaveMat layout (core/utils.h:359-373), read back the way misc/show_2Dmap.py:16-21 does."""
    from visma_b200 import io2d
    rng = np.random.default_rng(1)
    for arr in (rng.random((7, 5)).astype(np.float32), rng.integers(0, 255, (4, 9)).astype(np.uint8)):
        f = tmp_path / "m.bin"
        io2d.SaveMat(str(f), arr)
        raw = open(f, "rb").read()
        h, w = np.frombuffer(raw[:8], np.int32)
        assert (h, w) == arr.shape and len(raw) == 8 + arr.nbytes
        assert (np.frombuffer(raw[8:], arr.dtype).reshape(h, w) == arr).all()
        assert (io2d.LoadMat(str(f), arr.dtype) == arr).all()


@pytest.mark.gpu
def test_render_depth_tool(vb, oracle, tmp_path):
    """scripts/render_depth.py on the reference's own config values (misc/render_depth.json)."""
    import json
    import subprocess
    import sys
    from conftest import ROOT
    from visma_b200 import io2d
    cfg = {"major_version": 3, "minor_version": 3, "fx": 400, "fy": 400, "z_far": 10, "mesh": "misc/hermanmiller_aeron.obj",
           "translation": [0, 0, 1], "show": False, "save": True, "output_path": str(tmp_path), "mask": True}
    p = tmp_path / "render_depth.json"
    p.write_text("// comments are legal in the reference's configs\n" + json.dumps(cfg))
    out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "render_depth.py"), str(p)],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    depth = io2d.LoadMat(str(tmp_path / "depthmap.bin"), np.float32)
    mask = io2d.LoadMat(str(tmp_path / "mask.bin"), np.uint8)
    V, F = vb.synth.load_chair()
    P = oracle.projection(0.05, 10.0, 400.0, 400.0, 320.0, 400.0, 480, 640)   # cy := fy, render_depth.cpp:31
    oz, od = oracle.render_depth(V, F, vb.synth.make_T(np.eye(3), [0, 0, 1.0]).astype(np.float32).T.reshape(-1),
                                 oracle.view(np.eye(4, dtype=np.float32).reshape(-1)), P, 480, 640)
    assert depth.shape == (480, 640) and (depth == od).all()
    assert (mask == oracle.render_mask(oz)).all()
