"""GPU parity: radius-bounded 1-NN (vb200_knn1) against the oracle — indices AND double distances must be
bit-identical (integer/index work).  Run on the B200 box: pytest -m gpu."""
import numpy as np
import pytest

from conftest import small_scene

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def scene():
    return small_scene(n_scene=120000, n_objects=3, m=3000)


def test_knn_matches_oracle_bitexact(vb, oracle, scene):
    tgt = scene["scene_xyz"]
    q = np.concatenate([
        vb.synth.knn_queries(tgt, 20000, sigma=0.01),
        vb.synth.knn_queries(tgt, 5000, sigma=0.08),        # many beyond the radius
        tgt[:2000],                                          # exactly on scene points: d2 == 0
        np.random.default_rng(1).uniform(-5, 12, (2000, 3)),  # far outside the room / the grid
    ])
    sc = vb.reg.Scene(tgt, 0.075)
    gi, gd = sc.SearchHybrid1(q, 0.075)
    oi, od = oracle.Index(tgt, 0.075).knn1(q, 0.075)
    assert (gi == oi).all()
    assert (gd == od).all()
    assert (gi >= 0).sum() > 20000 and (gi < 0).sum() > 2000
    # a smaller radius on the same grid
    gi, gd = sc.SearchHybrid1(q, 0.02)
    oi, od = oracle.Index(tgt, 0.02).knn1(q, 0.02)
    assert (gi == oi).all() and (gd == od).all()


def test_threshold_and_ties(vb, oracle):
    r = 0.075
    r2f = float(np.float32(r * r))
    d_in = np.sqrt(min(r2f, r * r)) * (1 - 1e-9)
    d_between = np.sqrt((r2f + r * r) / 2)
    d_out = np.sqrt(max(r2f, r * r)) * (1 + 1e-9)
    # duplicates and mirrored points give exact ties -> lowest target index wins
    tgt = np.array([[0, 0, 0.0], [0, 0, 0.0], [1.0, 0, 0], [1.02, 0, 0], [1.0, 0, 0], [5, 5, 5]])
    q = np.array([[d_in, 0, 0], [d_between, 0, 0], [d_out, 0, 0], [1.01, 0, 0], [1.0, 0, 0], [0, 0, 0.0]])
    sc = vb.reg.Scene(tgt, r)
    gi, gd = sc.SearchHybrid1(q, r)
    oi, od = oracle.Index(tgt, r).knn1(q, r)
    assert (gi == oi).all() and (gd == od).all()
    assert gi[0] == 0 and gi[2] == -1 and gi[3] == 2 and gi[4] == 2 and gi[5] == 0


def test_near_ties_resolved_in_double(vb, oracle):
    """Pairs of targets whose distances to the query differ by ~1e-12 relative: float screening cannot
    separate them, the double re-check must."""
    rng = np.random.default_rng(2)
    base = rng.uniform(0, 4, (3000, 3))
    dirs = rng.normal(size=(3000, 3))
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    d = rng.uniform(0.005, 0.05, (3000, 1))
    a = base + dirs * d
    b = base - dirs * d * (1 + 1e-12 * rng.choice([-1, 1], (3000, 1)))
    tgt = np.concatenate([a, b])
    sc = vb.reg.Scene(tgt, 0.075)
    gi, gd = sc.SearchHybrid1(base, 0.075)
    oi, od = oracle.Index(tgt, 0.075).knn1(base, 0.075)
    assert (gi == oi).all() and (gd == od).all()
    # the exhaustive search screens in f32 too (from 16 queries up): same answers, and with the screen switched off
    bi, bd = vb.reg.SearchHybrid1BruteForce(tgt, base, 0.075)
    assert (bi == oi).all() and (bd == od).all()


def test_docs_kat_correspondences(vb, oracle, kat):
    s, t = kat["src"].astype(np.float64), kat["tgt"].astype(np.float64)
    T = kat["init"]
    q = s @ T[:3, :3].T + T[:3, 3]
    sc = vb.reg.Scene(t, 0.02)
    gi, gd = sc.SearchHybrid1(q, 0.02)
    oi, od = oracle.Index(t, 0.02).knn1(q, 0.02)
    assert (gi == oi).all() and (gd == od).all()
    # 34 741 correspondences is what the docs publish for this init (icp_registration.rst:56-58); the
    # transform above is numpy's, so allow the count to move by a couple of threshold cases
    assert abs(int((gi >= 0).sum()) - 34741) <= 2


def test_edge_cases(vb):
    one = np.array([[1.0, 2.0, 3.0]])
    sc = vb.reg.Scene(one, 0.1)
    assert sc.size()["n"] == 1
    i, d = sc.SearchHybrid1(np.zeros((0, 3)), 0.1)
    assert len(i) == 0
    i, d = sc.SearchHybrid1(np.array([[1.0, 2.0, 3.05], [1.0, 2.0, 3.2], [np.nan, 0, 0]]), 0.1)
    assert list(i) == [0, -1, -1] and d[1] == 0.0
    with pytest.raises(vb.pkg.VismaB200Error):   # radius beyond the grid's cell
        sc.SearchHybrid1(one, 0.5)
    with pytest.raises(vb.pkg.VismaB200Error):
        sc.SearchHybrid1(one, 0.0)


def test_full_size_properties(vb):
    """BASELINE size (2 M-pt scene): size-independent properties instead of an oracle run."""
    d = vb.synth.make_room_scene(2_000_000, 8, 100)
    tgt = d["scene_xyz"]
    sc = vb.reg.Scene(tgt, 0.075)
    info = sc.size()
    assert info["n"] == 2_000_000 and info["fine_cells"] > 0
    # every scene point finds a neighbour at distance 0 (itself, or a lower-index duplicate)
    sel = np.random.default_rng(0).choice(len(tgt), 200000, replace=False)
    i, d2 = sc.SearchHybrid1(tgt[sel], 0.075)
    assert (d2 == 0).all() and (i <= sel).all() and (tgt[i] == tgt[sel]).all()
    # reported distance is the distance to the reported point, and no sampled scene point is closer
    q = vb.synth.knn_queries(tgt, 20000, sigma=0.02)
    i, d2 = sc.SearchHybrid1(q, 0.075)
    m = i >= 0
    dd = q[m] - tgt[i[m]]
    assert (((dd[:, 0] ** 2 + dd[:, 1] ** 2) + dd[:, 2] ** 2) == d2[m]).all()
    probe = tgt[np.random.default_rng(1).choice(len(tgt), 2000, replace=False)]
    dm = ((q[:, None, :] - probe[None, :, :]) ** 2).sum(-1).min(1)
    r2 = float(np.float32(0.075 * 0.075))
    assert (np.where(m, d2, r2) <= dm + 1e-15).all()


def test_large_coherent_batch_path(vb, oracle, scene):
    """More than 262 144 queries take the grid-ordered, warp-cooperative kernel instead of warp-per-query."""
    tgt = scene["scene_xyz"]
    q = np.concatenate([vb.synth.knn_queries(tgt, 280000, sigma=0.01, seed=3),
                        vb.synth.knn_queries(tgt, 20000, sigma=0.1, seed=4)])
    sc = vb.reg.Scene(tgt, 0.075)
    gi, gd = sc.SearchHybrid1(q, 0.075)
    oi, od = oracle.Index(tgt, 0.075).knn1(q, 0.075)
    assert (gi == oi).all() and (gd == od).all()


def test_bruteforce_matches_oracle_bitexact(vb, oracle):
    """vb200_knn1_bruteforce (TMA-staged exhaustive search) against the oracle: ragged tile / chunk sizes,
    targets smaller than one tile, none at all, ties."""
    rng = np.random.default_rng(7)
    for n, nq in ((5000, 301), (1024, 8), (1023, 7), (3 * 1024 + 1, 64), (17, 5), (9000, 1), (2500, 2), (2500, 3), (4100, 4)):
        tgt = rng.uniform(0, 1, (n, 3))
        q = np.concatenate([tgt[: nq // 2] + rng.normal(0, 0.01, (nq // 2, 3)), rng.uniform(-1, 2, (nq - nq // 2, 3))])
        gi, gd = vb.reg.SearchHybrid1BruteForce(tgt, q, 0.075)
        oi, od = oracle.Index(tgt, 0.075).knn1(q, 0.075)
        assert (gi == oi).all() and (gd == od).all(), (n, nq)
    # duplicates: exact ties go to the lowest target index; empty target; empty query set
    tgt = np.array([[0, 0, 0.0], [0, 0, 0.0], [1.0, 0, 0], [1.02, 0, 0], [1.0, 0, 0], [5, 5, 5]])
    q = np.array([[1.01, 0, 0], [1.0, 0, 0], [0, 0, 0.0], [3, 3, 3.0]])
    gi, gd = vb.reg.SearchHybrid1BruteForce(tgt, q, 0.075)
    assert list(gi) == [2, 2, 0, -1] and gd[3] == 0.0
    # the same through the f32-screened kernel (16 queries or more), with the acceptance threshold's edge cases
    r2f = float(np.float32(0.075 * 0.075))
    d_in, d_out = np.sqrt(min(r2f, 0.075 * 0.075)) * (1 - 1e-9), np.sqrt(max(r2f, 0.075 * 0.075)) * (1 + 1e-9)
    q16 = np.concatenate([q, [[d_in, 0, 0], [np.sqrt((r2f + 0.075 * 0.075) / 2), 0, 0], [d_out, 0, 0]], np.tile(q, (4, 1))])
    gi, gd = vb.reg.SearchHybrid1BruteForce(tgt, q16, 0.075)
    oi, od = oracle.Index(tgt, 0.075).knn1(q16, 0.075)
    assert (gi == oi).all() and (gd == od).all() and list(gi[:4]) == [2, 2, 0, -1] and gi[4] == 0 and gi[6] == -1
    # far from the origin: the screen centres the coordinates (a cloud 1 km away keeps millimetre structure in f32)
    far = np.array([1000.0, -2000.0, 500.0])
    tg2 = rng.uniform(0, 1, (6000, 3)) + far
    q2 = tg2[:400] + rng.normal(0, 0.003, (400, 3))
    gi, gd = vb.reg.SearchHybrid1BruteForce(tg2, q2, 0.075)
    oi, od = oracle.Index(tg2, 0.075).knn1(q2, 0.075)
    assert (gi == oi).all() and (gd == od).all()
    gi, gd = vb.reg.SearchHybrid1BruteForce(np.zeros((0, 3)), q, 0.075)
    assert (gi == -1).all() and (gd == 0).all()
    gi, gd = vb.reg.SearchHybrid1BruteForce(tgt, np.zeros((0, 3)), 0.075)
    assert len(gi) == 0
    with pytest.raises(vb.pkg.VismaB200Error):
        vb.reg.SearchHybrid1BruteForce(tgt, q, 0.0)


def test_grid_search_equals_exhaustive_search_at_full_size(vb):
    """BASELINE size: the grid search and the index-free exhaustive search are two independent implementations of
    the same operator; on the 2 M-point scene they must agree bit for bit (indices and double distances) —
    2e10 distance evaluations no CPU oracle finishes in a test."""
    d = vb.synth.make_room_scene(2_000_000, 8, 100)
    tgt = d["scene_xyz"]
    q = np.concatenate([vb.synth.knn_queries(tgt, 8000, sigma=0.01), vb.synth.knn_queries(tgt, 2000, sigma=0.08, seed=9)])
    sc = vb.reg.Scene(tgt, 0.075)
    gi, gd = sc.SearchHybrid1(q, 0.075)
    bi, bd = vb.reg.SearchHybrid1BruteForce(tgt, q, 0.075)
    assert (gi == bi).all() and (gd == bd).all()
    assert (gi >= 0).sum() > 7000 and (gi < 0).sum() > 300


def test_grid_search_equals_exhaustive_search_at_ten_million_points(vb, oracle):
    """The largest size of BASELINE's KNN sweep (config 5: 10 M scene points x 10 000 queries): grid search vs the
    exhaustive search bit for bit on all 10 000 queries, and both against the CPU oracle's exhaustive scan on 48 of
    them (4.8e8 double distance evaluations)."""
    d = vb.synth.make_room_scene(10_000_000, 8, 10)
    tgt = d["scene_xyz"]
    q = np.concatenate([vb.synth.knn_queries(tgt, 9000, sigma=0.01), vb.synth.knn_queries(tgt, 1000, sigma=0.08, seed=9)])
    sc = vb.reg.Scene(tgt, 0.075)
    gi, gd = sc.SearchHybrid1(q, 0.075)
    bi, bd = vb.reg.SearchHybrid1BruteForce(tgt, q, 0.075)
    assert (gi == bi).all() and (gd == bd).all()
    assert (gi >= 0).sum() > 9000 and (gi < 0).sum() > 100
    pick = np.r_[0:24, 9000:9024]
    oi, od = oracle.knn1_brute(tgt, q[pick])
    r2 = float(np.float32(0.075 * 0.075))
    hit = od < r2
    assert np.array_equal(np.where(hit, oi, -1), gi[pick]) and np.array_equal(np.where(hit, od, 0.0), gd[pick])


def test_release_cached_memory_leaves_live_objects_alone(vb, oracle):
    """vb200_release_cached_memory trims the stream-ordered pool the library allocates from: a live scene keeps
    answering, and the next scene is built as before."""
    rng = np.random.default_rng(4)
    tgt = rng.uniform(0, 1, (20000, 3))
    q = tgt[:500] + rng.normal(0, 0.004, (500, 3))
    sc = vb.reg.Scene(tgt, 0.05)
    a = sc.SearchHybrid1(q, 0.05)
    vb.reg.Scene(tgt[:5000], 0.05).close()  # something freed into the pool
    assert vb.lib.lib().vb200_release_cached_memory(0) == 0
    b = sc.SearchHybrid1(q, 0.05)
    c = vb.reg.Scene(tgt, 0.05).SearchHybrid1(q, 0.05)
    oi, od = oracle.Index(tgt, 0.05).knn1(q, 0.05)
    for gi, gd in (a, b, c):
        assert (gi == oi).all() and (gd == od).all()


def test_nan_points_are_never_neighbours(vb, oracle):
    """A NaN target or query point matches nothing (the reference's `dist < worst_dist` is false for NaN), in the
    grid search and in the exhaustive one, whatever the NaN's sign bit."""
    rng = np.random.default_rng(11)
    tgt = rng.uniform(0, 1, (3000, 3))
    tgt[5] = np.nan
    tgt[700, 1] = -np.nan
    tgt[1500, 2] = np.copysign(np.nan, -1.0)
    q = np.concatenate([tgt[:1000] + 0.001, [[np.nan, 0.5, 0.5], [0.5, np.copysign(np.nan, -1.0), 0.5]]])
    oi, od = oracle.Index(tgt, 0.075).knn1(q, 0.075)
    gi, gd = vb.reg.Scene(tgt, 0.075).SearchHybrid1(q, 0.075)
    bi, bd = vb.reg.SearchHybrid1BruteForce(tgt, q, 0.075)
    assert (gi == oi).all() and (gd == od).all()
    assert (bi == oi).all() and (bd == od).all()
    assert not np.isin(gi, [5, 700, 1500]).any() and gi[-1] == -1 and gi[-2] == -1


def test_pageable_upload_is_staged_and_exact(vb):
    """Uploads of 16 MB or more from PAGEABLE host memory go through the library's staged copy (host threads ->
    pinned slots -> copy engine, runtime.cu h2d_async): the scene built from a pageable array — odd size, source
    pointer off the 16-byte grid — must answer exactly like the scene built from a pinned copy of the same points
    (which takes the direct path), and like the exhaustive search."""
    import torch
    n = 1_000_003                                    # 24 MB of coordinates + 24 MB of normals, not a multiple of any chunk
    d = vb.synth.make_room_scene(n, 1, 1000, seed=11)
    raw = np.empty(3 * n + 1, np.float64)
    xyz = raw[1:].reshape(n, 3)                      # 8 bytes off the allocation's alignment
    xyz[:] = d["scene_xyz"]
    nrm = np.ascontiguousarray(d["scene_nrm"])
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    q = xyz[::211] + 0.004
    a = vb.reg.Scene(vb.reg.PointCloud(xyz, nrm), 0.075)
    b = vb.reg.Scene(vb.reg.PointCloud(pin(xyz), pin(nrm)), 0.075)
    ia, da = a.SearchHybrid1(q, 0.075)
    ib, db = b.SearchHybrid1(q, 0.075)
    assert (ia == ib).all() and (da == db).all() and (ia >= 0).sum() > len(q) // 2
    # the exhaustive search uploads the same pageable array itself (staged as well)
    ic, dc = vb.reg.SearchHybrid1BruteForce(xyz, q[:64], 0.075)
    assert (ic == ia[:64]).all() and (dc == da[:64]).all()
    # and the normals went up intact: a point-to-plane alignment of a slice of the scene onto itself is the identity
    src = vb.reg.PointCloud(xyz[5000:25000].copy(), nrm[5000:25000].copy())
    for s in (a, b):
        r = vb.reg.RegistrationICP(src, s, 0.02, np.eye(4), vb.reg.TransformationEstimationPointToPlane())
        assert r.fitness_ == 1.0 and np.abs(r.transformation_ - np.eye(4)).max() < 1e-12
    a.close(); b.close()
