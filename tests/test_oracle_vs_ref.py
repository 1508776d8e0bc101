"""The plain-C restatement against the UNMODIFIED reference (oracle/_ref, built from /root/reference) on
seeded synthetic inputs — covers what the reference's own tests do not pin (estimators, the cicp class,
RegisterModelToScene, VoxelDownSample at size).  CPU only; skipped where oracle/_ref is not built."""
import numpy as np
import pytest

from conftest import small_scene


@pytest.fixture(scope="module")
def scene():
    return small_scene()


def test_knn_bitexact(oracle, ref, scene):
    tgt = scene["scene_xyz"]
    from visma_b200 import synth
    q = synth.knn_queries(tgt, 5000)
    kd = ref.KDTree(tgt)
    ri, rd = kd.search_hybrid1(q, 0.075)
    oi, od = oracle.Index(tgt, 0.075).knn1(q, 0.075)
    assert (ri == oi).all() and (rd == od).all()
    assert (oi >= 0).sum() > 4000


def test_threshold_is_float_rounded(oracle, ref):
    # KDTreeFlann.cpp:185 hands FLANN float(radius*radius): a neighbour at d2 between the double and the
    # float-rounded threshold decides which one is in force
    r = 0.075
    r2f = float(np.float32(r * r))
    assert r2f != r * r
    d_in = np.sqrt(min(r2f, r * r)) * (1 - 1e-9)
    d_between = np.sqrt((r2f + r * r) / 2)
    tgt = np.array([[0, 0, 0.0], [10, 10, 10]])
    q = np.array([[d_in, 0, 0], [d_between, 0, 0]])
    ri, rd = ref.KDTree(tgt).search_hybrid1(q, r)
    oi, od = oracle.Index(tgt, r).knn1(q, r)
    assert (ri == oi).all() and (rd == od).all()
    assert oi[0] == 0 and (oi[1] == 0) == (d_between ** 2 < r2f)


@pytest.mark.parametrize("est", ["p2p", "p2plane", "cicp"])
def test_estimators(oracle, ref, scene, est):
    tgt, tn = scene["scene_xyz"], scene["scene_nrm"]
    src = scene["sources"][0][0] @ scene["T_init"][0][:3, :3].T + scene["T_init"][0][:3, 3]
    oi, _ = oracle.Index(tgt, 0.075).knn1(src, 0.075)
    corr = np.stack([np.nonzero(oi >= 0)[0], oi[oi >= 0]], 1).astype(np.int32)
    assert len(corr) > 1000
    o_kind = {"p2p": oracle.P2P, "p2plane": oracle.P2PLANE, "cicp": oracle.P2P_CICP}[est]
    r_kind = {"p2p": ref.P2P, "p2plane": ref.P2PLANE, "cicp": ref.CICP_4DOF}[est]
    To = oracle.estimate(src, tgt, corr, o_kind, tgt_nrm=tn)
    Tr = ref.estimate(src, tgt, corr, r_kind, tgt_nrm=tn)
    assert np.allclose(To, Tr, atol=1e-11)
    assert np.allclose(To[:3, :3] @ To[:3, :3].T, np.eye(3), atol=1e-12)


@pytest.mark.parametrize("est", ["p2p", "p2plane"])
def test_icp_loop(oracle, ref, scene, est):
    tgt, tn = scene["scene_xyz"], scene["scene_nrm"]
    ix = oracle.Index(tgt, 0.075)
    for b in range(2):
        src, sn = scene["sources"][b]
        o = ix.registration_icp(src, 0.075, scene["T_init"][b], oracle.P2P if est == "p2p" else oracle.P2PLANE,
                                src_nrm=sn, tgt_nrm=tn, want_corr=True)
        r = ref.registration_icp(src, tgt, 0.075, scene["T_init"][b], ref.P2P if est == "p2p" else ref.P2PLANE,
                                 src_nrm=sn, tgt_nrm=tn, want_corr=True)
        assert o["ncorr"] == r["ncorr"] and o["fitness"] == r["fitness"]
        assert abs(o["rmse"] - r["rmse"]) < 1e-12
        assert np.allclose(o["T"], r["T"], atol=1e-10)
        assert (o["corr"] == r["corr"]).all()


def test_register_model_to_scene(oracle, ref):
    d = small_scene(n_scene=30000, n_objects=1, m=1500, seed=9)
    # the annotation tool works on a cropped scan: keep the scene points near the object
    c = d["T_gt"][0][:3, 3]
    scan = d["scene_xyz"][np.linalg.norm(d["scene_xyz"][:, [0, 2]] - c[[0, 2]], axis=1) < 0.8]
    model = d["sources"][0][0] @ d["T_gt"][0][:3, :3].T + c  # start at the true pose; yaw inits rotate about +Y
    o = oracle.register_model_to_scene(model, scan, level=6, threshold=0.05)
    r = ref.register_model_to_scene(model, scan, level=6, threshold=0.05)
    assert o["best_level"] == r["best_level"] and o["ncorr"] == r["ncorr"]
    assert np.allclose(o["T"], r["T"], atol=1e-10)


def test_voxel_downsample(oracle, ref, scene):
    xyz, nrm = scene["scene_xyz"][:20000], scene["scene_nrm"][:20000]
    o_p, o_n = oracle.voxel_downsample(xyz, 0.05, nrm)
    r_p, r_n = ref.voxel_downsample(xyz, 0.05, nrm)
    assert len(o_p) == len(r_p)
    ko = np.lexsort((o_p[:, 0], o_p[:, 1], o_p[:, 2]))
    kr = np.lexsort((r_p[:, 0], r_p[:, 1], r_p[:, 2]))
    assert (o_p[ko] == r_p[kr]).all()      # same summation order -> bit-identical averages
    assert np.allclose(o_n[ko], r_n[kr], atol=1e-15)


def test_transform(oracle, ref, scene):
    src, sn = scene["sources"][0]
    p, n = ref.transform(src, scene["T_init"][0], sn)
    T = scene["T_init"][0]
    assert np.allclose(p, src @ T[:3, :3].T + T[:3, 3], atol=1e-14)
    assert np.allclose(n, sn @ T[:3, :3].T, atol=1e-14)
