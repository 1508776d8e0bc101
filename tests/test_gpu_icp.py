"""GPU parity: the batched ICP operator, the estimator plug-in and the orientation-constrained driver
against the oracle and the reference's published known answers.  Run on the B200 box: pytest -m gpu.

Tolerances: BASELINE.json's north_star asks for converged poses within 1e-4 rad / 1e-3 m of the reference
CPU path; correspondences and counts are index work and must be identical."""
import numpy as np
import pytest

from conftest import small_scene

pytestmark = pytest.mark.gpu

ROT_TOL, TRANS_TOL = 1e-4, 1e-3


@pytest.fixture(scope="module")
def scene():
    return small_scene(n_scene=150000, n_objects=4, m=6000)


def clouds(vb, scene):
    return [vb.reg.PointCloud(p, n) for p, n in scene["sources"]]


def est_of(vb, oracle, name):
    return {"p2p": (vb.reg.TransformationEstimationPointToPoint(), oracle.P2P),
            "cicp": (vb.reg.TransformationEstimationPointToPoint4DoF(), oracle.P2P_CICP),
            "p2plane": (vb.reg.TransformationEstimationPointToPlane(), oracle.P2PLANE),
            "gravity": (vb.reg.TransformationEstimationPointToPlaneGravity((0, 1, 0)), oracle.P2PLANE_GRAVITY)}[name]


def test_first_pass_correspondences_identical(vb, oracle, scene):
    """max_iteration = 0 is EvaluateRegistration: same pairs, same fitness, rmse to rounding."""
    sc = vb.reg.Scene(vb.reg.PointCloud(scene["scene_xyz"], scene["scene_nrm"]), 0.075)
    res = vb.reg.RegistrationICPBatch(clouds(vb, scene), sc, 0.075, scene["T_init"],
                                      criteria=vb.reg.ICPConvergenceCriteria(1e-6, 1e-6, 0))
    ix = oracle.Index(scene["scene_xyz"], 0.075)
    for b, r in enumerate(res):
        o = ix.registration_icp(scene["sources"][b][0], 0.075, scene["T_init"][b], oracle.P2P, max_iter=0,
                                want_corr=True)
        assert r.iterations_ == 0 and np.array_equal(r.transformation_, scene["T_init"][b])
        assert len(r.correspondence_set_) == o["ncorr"] and (r.correspondence_set_ == o["corr"]).all()
        assert r.fitness_ == o["fitness"] and abs(r.inlier_rmse_ - o["rmse"]) < 1e-12


@pytest.mark.parametrize("name", ["p2p", "cicp", "p2plane", "gravity"])
def test_icp_batch_matches_oracle(vb, oracle, scene, name):
    est, okind = est_of(vb, oracle, name)
    sc = vb.reg.Scene(vb.reg.PointCloud(scene["scene_xyz"], scene["scene_nrm"]), 0.075)
    res = vb.reg.RegistrationICPBatch(clouds(vb, scene), sc, 0.075, scene["T_init"], est)
    ix = oracle.Index(scene["scene_xyz"], 0.075)
    for b, r in enumerate(res):
        src, sn = scene["sources"][b]
        o = ix.registration_icp(src, 0.075, scene["T_init"][b], okind, src_nrm=sn, tgt_nrm=scene["scene_nrm"],
                                want_corr=True)
        rot, tr = vb.synth.pose_error(r.transformation_, o["T"])
        assert rot < ROT_TOL and tr < TRANS_TOL, (name, b, rot, tr)
        # in practice the agreement is ~1e-9; keep an eye on it without making it the contract
        assert rot < 1e-6 and tr < 1e-6, (name, b, rot, tr)
        assert abs(r.fitness_ - o["fitness"]) <= 2.0 / len(src)
        assert abs(r.inlier_rmse_ - o["rmse"]) < 1e-6
        assert abs(r.iterations_ - o["iters"]) <= 1
        # converged pose is near the ground truth the scene was built from
        grot, gtr = vb.synth.pose_error(r.transformation_, scene["T_gt"][b])
        # (the 4-DoF estimator cannot undo the +-1 degree roll/pitch of T_init: it has its own test below)
        if name == "p2plane":
            assert grot < 5e-3 and gtr < 5e-3, (name, b, grot, gtr)


def test_gravity_estimator_keeps_gravity(vb, scene):
    """The 4-DoF estimator never rotates about anything but the gravity axis."""
    sc = vb.reg.Scene(vb.reg.PointCloud(scene["scene_xyz"], scene["scene_nrm"]), 0.075)
    inits = scene["T_gt"].copy()
    inits[:, :3, 3] += [0.02, 0.0, -0.015]
    res = vb.reg.RegistrationICPBatch(clouds(vb, scene), sc, 0.075, inits,
                                      vb.reg.TransformationEstimationPointToPlaneGravity((0, 1, 0)))
    for b, r in enumerate(res):
        d = r.transformation_ @ np.linalg.inv(inits[b])
        assert abs(d[1, 1] - 1) < 1e-12 and abs(d[0, 1]) < 1e-12 and abs(d[2, 1]) < 1e-12
        grot, gtr = vb.synth.pose_error(r.transformation_, scene["T_gt"][b])
        assert grot < 5e-3 and gtr < 5e-3


def test_docs_known_answer(vb, kat):
    """Open3D's published registration_icp outputs (docs/tutorial/Basic/icp_registration.rst:56-58,
    91-98,154-161), reproduced on the GPU."""
    s, t, tn = (kat[k].astype(np.float64) for k in ("src", "tgt", "tgt_nrm"))
    sc = vb.reg.Scene(vb.reg.PointCloud(t, tn), 0.02)
    src = vb.reg.PointCloud(s, s)  # normals: presence only
    ev = vb.reg.EvaluateRegistration(src, sc, 0.02, kat["init"])
    assert abs(ev.fitness_ - 0.174723) < 5e-7 and abs(ev.inlier_rmse_ - 0.011771) < 5e-7
    assert len(ev.correspondence_set_) == 34741
    r = vb.reg.RegistrationICP(src, sc, 0.02, kat["init"], vb.reg.TransformationEstimationPointToPoint())
    assert abs(r.fitness_ - 0.372450) < 5e-7 and abs(r.inlier_rmse_ - 0.007760) < 5e-7
    assert len(r.correspondence_set_) == 74056
    assert np.allclose(r.transformation_, kat["doc_p2p_T"], atol=5e-9)
    assert np.allclose(r.transformation_, kat["ref_p2p_T"], atol=1e-9)
    r = vb.reg.RegistrationICP(src, sc, 0.02, kat["init"], vb.reg.TransformationEstimationPointToPlane())
    assert abs(r.fitness_ - 0.620972) < 5e-7 and abs(r.inlier_rmse_ - 0.006581) < 5e-7
    assert len(r.correspondence_set_) == 123471
    assert np.allclose(r.transformation_, kat["doc_p2l_T"], atol=5e-9)
    assert np.allclose(r.transformation_, kat["ref_p2l_T"], atol=1e-9)


@pytest.mark.parametrize("name", ["p2p", "p2plane", "gravity"])
def test_estimator_plugin(vb, oracle, scene, name):
    """TransformationEstimation::ComputeTransformation on the GPU, driven by an explicit correspondence set
    (how the reference's CPU loop would call it)."""
    est, okind = est_of(vb, oracle, name)
    tgt, tn = scene["scene_xyz"], scene["scene_nrm"]
    T0 = scene["T_init"][1]
    src = scene["sources"][1][0] @ T0[:3, :3].T + T0[:3, 3]
    oi, _ = oracle.Index(tgt, 0.075).knn1(src, 0.075)
    corr = np.stack([np.nonzero(oi >= 0)[0], oi[oi >= 0]], 1).astype(np.int32)
    Tg = vb.reg.ComputeTransformation(est, vb.reg.PointCloud(src), vb.reg.PointCloud(tgt, tn), corr)
    To = oracle.estimate(src, tgt, corr, okind, tgt_nrm=tn, gravity=(0, 1, 0))
    assert np.allclose(Tg, To, atol=1e-10)
    # empty set -> Identity; point-to-plane without target normals -> Identity
    assert np.array_equal(vb.reg.ComputeTransformation(est, src, vb.reg.PointCloud(tgt, tn), np.zeros((0, 2))), np.eye(4))
    if name != "p2p":
        assert np.array_equal(vb.reg.ComputeTransformation(est, src, vb.reg.PointCloud(tgt), corr), np.eye(4))
        # rank-deficient system (all normals equal) -> det guard -> Identity (Utility/Eigen.cpp:41-52)
        flat = np.tile([0.0, 1.0, 0.0], (len(tgt), 1))
        Tg = vb.reg.ComputeTransformation(est, src, vb.reg.PointCloud(tgt, flat), corr)
        To = oracle.estimate(src, tgt, corr, okind, tgt_nrm=flat, gravity=(0, 1, 0))
        assert np.allclose(Tg, To, atol=1e-9)


def test_reference_error_behaviour(vb, scene):
    sc = vb.reg.Scene(vb.reg.PointCloud(scene["scene_xyz"], scene["scene_nrm"]), 0.075)
    sc_nonrm = vb.reg.Scene(scene["scene_xyz"], 0.075)
    src = clouds(vb, scene)[:2]
    inits = scene["T_init"][:2]
    # invalid distance -> RegistrationResult(init)  (Registration.cpp:148-151)
    for r, T in zip(vb.reg.RegistrationICPBatch(src, sc, 0.0, inits), inits):
        assert np.array_equal(r.transformation_, T) and r.fitness_ == 0 and len(r.correspondence_set_) == 0
    # point-to-plane without normals on either side -> RegistrationResult(init)  (:152-157)
    p2l = vb.reg.TransformationEstimationPointToPlane()
    for r, T in zip(vb.reg.RegistrationICPBatch(src, sc_nonrm, 0.075, inits, p2l), inits):
        assert np.array_equal(r.transformation_, T) and r.fitness_ == 0
    bare = [vb.reg.PointCloud(c.points_) for c in src]
    for r, T in zip(vb.reg.RegistrationICPBatch(bare, sc, 0.075, inits, p2l), inits):
        assert np.array_equal(r.transformation_, T) and r.fitness_ == 0


def test_ragged_and_degenerate_batches(vb, oracle, scene):
    sc = vb.reg.Scene(vb.reg.PointCloud(scene["scene_xyz"], scene["scene_nrm"]), 0.075)
    far = vb.reg.PointCloud(scene["sources"][0][0] + 100.0, scene["sources"][0][1])  # nothing in range
    empty = vb.reg.PointCloud(np.zeros((0, 3)), np.zeros((0, 3)))
    tiny = vb.reg.PointCloud(scene["sources"][1][0][:7], scene["sources"][1][1][:7])
    big = clouds(vb, scene)[2]
    inits = np.stack([np.eye(4), np.eye(4), scene["T_init"][1], scene["T_init"][2]])
    res = vb.reg.RegistrationICPBatch([far, empty, tiny, big], sc, 0.075, inits,
                                      vb.reg.TransformationEstimationPointToPlane())
    assert res[0].fitness_ == 0 and res[0].inlier_rmse_ == 0 and np.array_equal(res[0].transformation_, np.eye(4))
    assert res[0].iterations_ == 1  # identity update, then "converged" (Registration.cpp:179-183)
    assert res[1].fitness_ == 0 and len(res[1].correspondence_set_) == 0
    ix = oracle.Index(scene["scene_xyz"], 0.075)
    o = ix.registration_icp(tiny.points_, 0.075, inits[2], oracle.P2PLANE, src_nrm=tiny.normals_,
                            tgt_nrm=scene["scene_nrm"])
    assert np.allclose(res[2].transformation_, o["T"], atol=1e-7) and res[2].fitness_ == o["fitness"]
    o = ix.registration_icp(big.points_, 0.075, inits[3], oracle.P2PLANE, src_nrm=big.normals_,
                            tgt_nrm=scene["scene_nrm"])
    rot, tr = vb.synth.pose_error(res[3].transformation_, o["T"])
    assert rot < 1e-6 and tr < 1e-6
    assert vb.reg.RegistrationICPBatch([], sc, 0.075, np.zeros((0, 4, 4))) == []


def test_deterministic_and_shard_invariant(vb, scene):
    """Same bits run to run, and the same bits whether an object is solved alone or inside a batch — the
    property the multi-GPU sharding relies on."""
    sc = vb.reg.Scene(vb.reg.PointCloud(scene["scene_xyz"], scene["scene_nrm"]), 0.075)
    cl = clouds(vb, scene)
    est = vb.reg.TransformationEstimationPointToPlane()
    a = vb.reg.RegistrationICPBatch(cl, sc, 0.075, scene["T_init"], est)
    b = vb.reg.RegistrationICPBatch(cl, sc, 0.075, scene["T_init"], est)
    for x, y in zip(a, b):
        assert np.array_equal(x.transformation_, y.transformation_) and x.inlier_rmse_ == y.inlier_rmse_
    for k in (0, 3):
        solo = vb.reg.RegistrationICPBatch([cl[k]], sc, 0.075, scene["T_init"][k:k + 1], est)[0]
        assert np.array_equal(solo.transformation_, a[k].transformation_)
        assert (solo.correspondence_set_ == a[k].correspondence_set_).all()


def test_register_model_to_scene(vb, oracle):
    """feh::RegisterModelToScene: 24 (here 8) yaw initialisations, arg-max correspondences."""
    d = small_scene(n_scene=40000, n_objects=1, m=2500, seed=9)
    c = d["T_gt"][0][:3, 3]
    keep = np.linalg.norm(d["scene_xyz"][:, [0, 2]] - c[[0, 2]], axis=1) < 0.8
    scan, scan_n = d["scene_xyz"][keep], d["scene_nrm"][keep]
    # annotation.cpp centres scan and model on the origin (T1/T2, :111-132) so yaw inits rotate in place
    origin = np.array([c[0], 0.0, c[2]])
    scan = scan - origin
    model = d["sources"][0][0] @ d["T_gt"][0][:3, :3].T
    model_n = d["sources"][0][1] @ d["T_gt"][0][:3, :3].T
    sc = vb.reg.Scene(vb.reg.PointCloud(scan, scan_n), 0.05)
    for p2l in (False, True):
        g = vb.reg.RegisterModelToScene(vb.reg.PointCloud(model, model_n), sc, 8, 0.05, p2l)
        o = oracle.register_model_to_scene(model, scan, level=8, threshold=0.05, point_to_plane=p2l,
                                           model_nrm=model_n, scan_nrm=scan_n)
        assert g["best_level"] == o["best_level"] == 0
        assert abs(g["ncorr"] - o["ncorr"]) <= 2
        rot, tr = vb.synth.pose_error(g["T"], o["T"])
        assert rot < ROT_TOL and tr < TRANS_TOL
        assert rot < 1e-6 and tr < 1e-6


def test_full_size_workload(vb, oracle):
    """BASELINE config 3 shape on one GPU: 2 M-pt scene x 32 objects x 50k points.  Checked through
    size-independent properties plus the oracle on ALL 32 objects."""
    d = vb.synth.make_room_scene(2_000_000, 32, 50_000)
    sc = vb.reg.Scene(vb.reg.PointCloud(d["scene_xyz"], d["scene_nrm"]), 0.075)
    cl = [vb.reg.PointCloud(p, n) for p, n in d["sources"]]
    res = vb.reg.RegistrationICPBatch(cl, sc, 0.075, d["T_init"], vb.reg.TransformationEstimationPointToPlane(),
                                      want_corr=False)
    errs = np.array([vb.synth.pose_error(r.transformation_, T) for r, T in zip(res, d["T_gt"])])
    assert (errs[:, 0] < 5e-3).all() and (errs[:, 1] < 5e-3).all(), errs.max(0)
    assert all(r.fitness_ > 0.95 for r in res)
    # every one of the 32 objects against the oracle (~1 s of CPU each)
    ix = oracle.Index(d["scene_xyz"], 0.075)
    for b in range(32):
        o = ix.registration_icp(d["sources"][b][0], 0.075, d["T_init"][b], oracle.P2PLANE,
                                src_nrm=d["sources"][b][1], tgt_nrm=d["scene_nrm"])
        rot, tr = vb.synth.pose_error(res[b].transformation_, o["T"])
        assert rot < ROT_TOL and tr < TRANS_TOL and rot < 1e-6 and tr < 1e-6, (b, rot, tr)
        assert abs(res[b].fitness_ - o["fitness"]) <= 2.0 / 50_000
        assert abs(res[b].inlier_rmse_ - o["rmse"]) < 1e-6 and abs(res[b].iterations_ - o["iters"]) <= 1


@pytest.mark.parametrize("name", ["p2p", "p2plane"])
def test_split_iteration_equals_fused(vb, scene, name):
    """vb200_batch_pass / (sum of totals) / vb200_batch_solve — the iteration split at the cross-GPU exchange
    point — on two shards of one cloud must reproduce the single-batch RegistrationICP of the whole cloud."""
    torch = pytest.importorskip("torch")
    est = (vb.reg.TransformationEstimationPointToPoint() if name == "p2p"
           else vb.reg.TransformationEstimationPointToPlane())
    sc = vb.reg.Scene(vb.reg.PointCloud(scene["scene_xyz"], scene["scene_nrm"]), 0.075)
    src, sn = scene["sources"][0]
    init = scene["T_init"][0]
    crit = vb.reg.ICPConvergenceCriteria()
    whole = vb.reg.RegistrationICP(vb.reg.PointCloud(src, sn), sc, 0.075, init, est, crit)
    cut = 2500
    shards = [vb.reg.PointCloud(src[:cut], sn[:cut]), vb.reg.PointCloud(src[cut:], sn[cut:])]
    batches, totals = [], []
    for sh in shards:
        b = vb.reg.Batch(sc, [sh])
        b.set_problems(init.reshape(1, 16))
        t = torch.zeros(32, dtype=torch.float64, device="cuda")
        b.set_totals_buffer(t.data_ptr())
        batches.append(b)
        totals.append(t)
    for it in range(crit.max_iteration_ + 1):
        for b in batches:
            b.pass_(est, 0.075)
        sc.sync()
        tot = totals[0] + totals[1]          # what the all-reduce produces on every rank
        torch.cuda.synchronize()
        for t in totals:
            t.copy_(tot)
        torch.cuda.synchronize()
        for b in batches:
            b.solve(est, 0.075, crit, it, [len(src)])
    ra, rb = batches[0].results()[0], batches[1].results()[0]
    assert np.array_equal(ra.transformation_, rb.transformation_)  # ranks stay consistent without a broadcast
    assert ra.fitness_ == rb.fitness_ and ra.iterations_ == rb.iterations_
    assert np.allclose(ra.transformation_, whole.transformation_, atol=1e-9, rtol=0)
    assert abs(ra.fitness_ - whole.fitness_) <= 1.0 / len(src) and abs(ra.inlier_rmse_ - whole.inlier_rmse_) < 1e-9
    assert ra.iterations_ == whole.iterations_


@pytest.mark.parametrize("name", ["p2p", "p2plane"])
def test_cached_neighbours_equal_fresh_search(vb, scene, name):
    """k_pass keeps, per source point, what its last search proved and skips the search while the triangle
    inequality shows the neighbour unchanged.  That must be invisible: after every iteration of a running
    batch the correspondences equal those of a FRESH batch (no history) evaluated at the same transforms,
    and they equal the independent warp-per-query KNN of the transformed points."""
    est = {"p2p": vb.reg.TransformationEstimationPointToPoint(),
           "p2plane": vb.reg.TransformationEstimationPointToPlane()}[name]
    sc = vb.reg.Scene(vb.reg.PointCloud(scene["scene_xyz"], scene["scene_nrm"]), 0.075)
    cl = clouds(vb, scene)
    run = vb.reg.Batch(sc, cl)
    run.set_problems(scene["T_init"])
    T = np.array(scene["T_init"], dtype=np.float64)
    for it in range(14):
        run.iterate(est, 0.075, 1)          # pass at T (cache from the previous iterations), then T <- update * T
        got = run.results(want_corr=True)
        fresh = vb.reg.Batch(sc, cl)
        fresh.set_problems(T)
        fresh.iterate(est, 0.075, 1)
        want = fresh.results(want_corr=True)
        for b, (g, w) in enumerate(zip(got, want)):
            assert np.array_equal(g.correspondence_set_, w.correspondence_set_), (it, b)
            assert np.allclose(g.transformation_, w.transformation_, rtol=1e-12, atol=1e-14)
            assert abs(g.inlier_rmse_ - w.inlier_rmse_) < 1e-13
        if it in (0, 5, 13):
            # the other search kernel (one warp per query, no history) on the same transformed points
            for b in (0, len(cl) - 1):
                q = cl[b].points_ @ T[b][:3, :3].T + T[b][:3, 3]
                idx, _ = sc.SearchHybrid1(q[:3000], 0.075)
                full = -np.ones(len(q), np.int64)
                cs = got[b].correspondence_set_
                full[cs[:, 0]] = cs[:, 1]
                assert np.array_equal(full[:3000], idx), (it, b)
        T = np.stack([r.transformation_ for r in got])
        fresh.close()


def test_cache_survives_set_problems_and_tiny_moves(vb, scene):
    """set_problems() forgets the history; a transform nudged by less than the search slack re-uses it.  Either
    way the answer is that of a fresh search."""
    est = vb.reg.TransformationEstimationPointToPlane()
    sc = vb.reg.Scene(vb.reg.PointCloud(scene["scene_xyz"], scene["scene_nrm"]), 0.075)
    cl = clouds(vb, scene)
    b1 = vb.reg.Batch(sc, cl)
    b1.set_problems(scene["T_init"])
    b1.iterate(est, 0.075, 12)
    T12 = np.stack([r.transformation_ for r in b1.results()])
    nudged = T12.copy()
    nudged[:, :3, 3] += 2e-4       # 0.2 mm: inside the slack, most points keep their cached neighbour
    for T in (T12, nudged, scene["T_init"]):
        b1.set_problems(T)
        # warm the history at T, then evaluate once more at the resulting transform
        b1.iterate(est, 0.075, 1)
        Tn = np.stack([r.transformation_ for r in b1.results()])
        b1.iterate(est, 0.075, 1)
        got = b1.results(want_corr=True)
        b2 = vb.reg.Batch(sc, cl)
        b2.set_problems(Tn)
        b2.iterate(est, 0.075, 1)
        want = b2.results(want_corr=True)
        for g, w in zip(got, want):
            assert np.array_equal(g.correspondence_set_, w.correspondence_set_)
        b2.close()


def test_exact_ties_through_the_cached_path(vb, scene):
    """Every scene point twice: each query's winner and runner-up are at EXACTLY the same distance, so every
    decision goes through the tie rule (lowest original index) — in the search, in the two-candidate cache
    entries and in part A's test.  The running batch must agree with the warp-per-query KNN at every iteration."""
    xyz = np.concatenate([scene["scene_xyz"][:40000], scene["scene_xyz"][:40000]])
    nrm = np.concatenate([scene["scene_nrm"][:40000], scene["scene_nrm"][:40000]])
    sc = vb.reg.Scene(vb.reg.PointCloud(xyz, nrm), 0.075)
    cl = clouds(vb, scene)[:2]
    est = vb.reg.TransformationEstimationPointToPlane()
    run = vb.reg.Batch(sc, cl)
    T = np.array(scene["T_init"][:2], dtype=np.float64)
    run.set_problems(T)
    for it in range(10):
        run.iterate(est, 0.075, 1)
        got = run.results(want_corr=True)
        for b in range(2):
            q = cl[b].points_ @ T[b][:3, :3].T + T[b][:3, 3]
            idx, _ = sc.SearchHybrid1(q, 0.075)
            full = -np.ones(len(q), np.int64)
            cs = got[b].correspondence_set_
            full[cs[:, 0]] = cs[:, 1]
            # the transform is applied with FMAs on the device and without in numpy: a query may differ in the
            # last bit, which can only matter for a point exactly on the radius — none here
            assert np.array_equal(full, idx), (it, b, int((full != idx).sum()))
            assert (idx[idx >= 0] < 40000).all()      # ties went to the first copy
        T = np.stack([r.transformation_ for r in got])


def test_compute_rmse(vb, oracle, scene):
    """A1: cicp::TransformationEstimationPointToPoint4DoF::ComputeRMSE (src/constrained_ICP.cpp:13-23) on the GPU."""
    tgt = scene["scene_xyz"]
    T0 = scene["T_init"][2]
    src = scene["sources"][2][0] @ T0[:3, :3].T + T0[:3, 3]
    oi, _ = oracle.Index(tgt, 0.075).knn1(src, 0.075)
    corr = np.stack([np.nonzero(oi >= 0)[0], oi[oi >= 0]], 1).astype(np.int32)
    want = np.sqrt(((src[corr[:, 0]] - tgt[corr[:, 1]]) ** 2).sum(1).sum() / len(corr))
    got = vb.reg.ComputeRMSE(src, tgt, corr)
    assert abs(got - want) < 1e-14 and got > 0
    assert vb.reg.ComputeRMSE(src, tgt, np.zeros((0, 2), np.int32)) == 0.0


@pytest.mark.parametrize("name", ["p2p", "p2plane", "gravity"])
def test_estimator_device_resident(vb, oracle, scene, name):
    """vb200_estimate_device: clouds and the correspondence list already on the GPU, rows gathered by the kernel."""
    torch = pytest.importorskip("torch")
    est, okind = est_of(vb, oracle, name)
    tgt, tn = scene["scene_xyz"], scene["scene_nrm"]
    T0 = scene["T_init"][1]
    src = scene["sources"][1][0] @ T0[:3, :3].T + T0[:3, 3]
    oi, _ = oracle.Index(tgt, 0.075).knn1(src, 0.075)
    corr = np.stack([np.nonzero(oi >= 0)[0], oi[oi >= 0]], 1).astype(np.int32)
    dev = [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (src, tgt, tn, corr)]
    torch.cuda.synchronize()
    Tg = vb.reg.ComputeTransformationDevice(est, dev[0].data_ptr(), len(src), dev[1].data_ptr(), dev[2].data_ptr(),
                                            len(tgt), dev[3].data_ptr(), len(corr))
    To = oracle.estimate(src, tgt, corr, okind, tgt_nrm=tn, gravity=(0, 1, 0))
    assert np.allclose(Tg, To, atol=1e-10)
    Th = vb.reg.ComputeTransformation(est, vb.reg.PointCloud(src), vb.reg.PointCloud(tgt, tn), corr)
    assert np.allclose(Tg, Th, atol=1e-12)
