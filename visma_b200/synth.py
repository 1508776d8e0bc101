"""Seeded synthetic workloads for the tests and bench.py: the real clutter1 data is a Dropbox
download the reference does not ship, so the scene, the object fragments and the render poses are generated.

Everything here is plain numpy on the host; it only PRODUCES inputs (it is not on the measured path).
"""
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def load_chair():
    """misc/hermanmiller_aeron.obj as (V float32 2492x3, F int32 4999x3) — see scripts/make_mesh_fixture.py."""
    d = np.load(os.path.join(_HERE, "data", "chair_mesh.npz"))
    return d["V"].astype(np.float32), d["F"].astype(np.int32)


def cube_mesh():
    """A unit cube, 8 vertices / 12 triangles (stands in for misc/cube.ply)."""
    V = np.array([[x, y, z] for z in (-0.5, 0.5) for y in (-0.5, 0.5) for x in (-0.5, 0.5)], np.float32)
    F = np.array([[0, 1, 3], [0, 3, 2], [4, 6, 7], [4, 7, 5], [0, 4, 5], [0, 5, 1],
                  [2, 3, 7], [2, 7, 6], [0, 2, 6], [0, 6, 4], [1, 5, 7], [1, 7, 3]], np.int32)
    return V, F


def sample_mesh(V, F, n, rng):
    """Area-weighted surface samples + face normals.  Same algorithm as feh::SamplePointCloudFromMesh
    (include/geometry.h:29-64: area CDF, uniform barycentric) with a seeded generator and correct face
    indexing (the reference seeds from the clock and is off by one on the face index)."""
    V = np.asarray(V, np.float64)
    a, b, c = V[F[:, 0]], V[F[:, 1]], V[F[:, 2]]
    cr = np.cross(b - a, c - a)
    area = 0.5 * np.linalg.norm(cr, axis=1)
    cdf = np.cumsum(area)
    f = np.searchsorted(cdf, rng.random(n) * cdf[-1], side="right").clip(0, len(F) - 1)
    r1, r2 = rng.random(n), rng.random(n)
    s = np.sqrt(r1)
    w0, w1, w2 = 1 - s, s * (1 - r2), s * r2
    pts = w0[:, None] * a[f] + w1[:, None] * b[f] + w2[:, None] * c[f]
    nrm = cr[f] / np.maximum(np.linalg.norm(cr[f], axis=1, keepdims=True), 1e-300)
    return pts, nrm


def rot_y(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])


def rot_xyz(rx, ry, rz):
    cx, sx, cy, sy, cz, sz = np.cos(rx), np.sin(rx), np.cos(ry), np.sin(ry), np.cos(rz), np.sin(rz)
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def make_T(R, t):
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = t
    return T


def make_room_scene(n_scene, n_objects, pts_per_object, seed=20260117, room=(6.0, 3.0, 6.0),
                    object_fraction=0.3, noise=0.002, source_seed=0):
    """SURVEY §8d workload.  Returns dict with
       scene_xyz, scene_nrm  (N x 3 f64): floor (y = 0), four walls and the objects' surfaces, N(0, noise)
                             along the normal; gravity is -Y as in VISMA (src/annotation.cpp:84)
       sources  list of (M x 3 f64 points, M x 3 normals) in each object's MODEL frame
       T_gt     B x 4 x 4 model -> scene ground truth
       T_init   B x 4 x 4 ground truth composed with a perturbation (yaw +-5 deg, roll/pitch +-1 deg, +-3 cm)
    source_seed changes only the source fragments and perturbations (ranks of a multi-GPU run share one
    scene but each aligns its own fragments).
    """
    rng = np.random.default_rng(seed)
    V, F = load_chair()
    V = V.astype(np.float64)
    V[:, 1] -= V[:, 1].min()  # chair stands on y = 0
    X, Y, Z = room
    n_obj_pts = int(n_scene * object_fraction) if n_objects > 0 else 0
    n_bg = n_scene - n_obj_pts
    # background: floor + 4 walls, sampled by area
    areas = np.array([X * Z, X * Y, X * Y, Z * Y, Z * Y])
    cnt = np.floor(n_bg * areas / areas.sum()).astype(int)
    cnt[0] += n_bg - cnt.sum()
    parts, norms = [], []
    u = rng.random((cnt[0], 2))
    parts.append(np.stack([u[:, 0] * X, np.zeros(cnt[0]), u[:, 1] * Z], 1)); norms.append(np.tile([0, 1.0, 0], (cnt[0], 1)))
    u = rng.random((cnt[1], 2))
    parts.append(np.stack([u[:, 0] * X, u[:, 1] * Y, np.zeros(cnt[1])], 1)); norms.append(np.tile([0, 0, 1.0], (cnt[1], 1)))
    u = rng.random((cnt[2], 2))
    parts.append(np.stack([u[:, 0] * X, u[:, 1] * Y, np.full(cnt[2], Z)], 1)); norms.append(np.tile([0, 0, -1.0], (cnt[2], 1)))
    u = rng.random((cnt[3], 2))
    parts.append(np.stack([np.zeros(cnt[3]), u[:, 1] * Y, u[:, 0] * Z], 1)); norms.append(np.tile([1.0, 0, 0], (cnt[3], 1)))
    u = rng.random((cnt[4], 2))
    parts.append(np.stack([np.full(cnt[4], X), u[:, 1] * Y, u[:, 0] * Z], 1)); norms.append(np.tile([-1.0, 0, 0], (cnt[4], 1)))
    # objects on a floor grid, random yaw, scale 0.8-1.2
    T_gt, T_init, sources = [], [], []
    gx = int(np.ceil(np.sqrt(max(n_objects, 1))))
    per = [n_obj_pts // n_objects + (1 if b < n_obj_pts % n_objects else 0) for b in range(n_objects)] if n_objects else []
    for b in range(n_objects):
        orng = np.random.default_rng(seed + 1 + b)
        scale = orng.uniform(0.8, 1.2)
        yaw = orng.uniform(0, 2 * np.pi)
        cx = (b % gx + 0.5) * X / gx + orng.uniform(-0.1, 0.1)
        cz = (b // gx + 0.5) * Z / gx + orng.uniform(-0.1, 0.1)
        Vs = V * scale
        p, nn = sample_mesh(Vs, F, per[b], orng)
        R = rot_y(yaw)
        parts.append(p @ R.T + [cx, 0.0, cz]); norms.append(nn @ R.T)
        # the source fragment is sampled from the scaled CAD model in its own frame (rigid ground truth)
        srng = np.random.default_rng([seed + 1 + b, 7919, source_seed])
        sp, sn = sample_mesh(Vs, F, pts_per_object, srng)
        sources.append((sp, sn))
        Tr = make_T(R, [cx, 0.0, cz])
        d = np.deg2rad(srng.uniform([-1, -5, -1], [1, 5, 1]))
        Tp = make_T(rot_xyz(*d), srng.uniform(-0.03, 0.03, 3))
        # perturb about the object's own position so a 5 degree yaw does not swing it across the room
        C = make_T(np.eye(3), [cx, 0.0, cz])
        T_gt.append(Tr)
        T_init.append(C @ Tp @ np.linalg.inv(C) @ Tr)
    xyz = np.concatenate(parts)
    nrm = np.concatenate(norms)
    xyz = xyz + nrm * rng.normal(0.0, noise, (len(xyz), 1))
    perm = rng.permutation(len(xyz))  # scans are not ordered by surface
    return dict(scene_xyz=np.ascontiguousarray(xyz[perm]), scene_nrm=np.ascontiguousarray(nrm[perm]),
                sources=sources, T_gt=np.asarray(T_gt).reshape(-1, 4, 4), T_init=np.asarray(T_init).reshape(-1, 4, 4))


def knn_queries(scene_xyz, q, seed=7, sigma=0.01):
    """Config 5 queries: scene samples + N(0, 1 cm)."""
    rng = np.random.default_rng(seed)
    idx = rng.integers(0, len(scene_xyz), q)
    return scene_xyz[idx] + rng.normal(0.0, sigma, (q, 3))


def render_poses(n, seed=11):
    """Config 4 model poses: yaw uniform, t = [+-0.5, +-0.3, 1..3]; n x 4 x 4 float32 (math layout)."""
    rng = np.random.default_rng(seed)
    out = np.zeros((n, 4, 4), np.float32)
    for i in range(n):
        out[i] = make_T(rot_y(rng.uniform(0, 2 * np.pi)),
                        [rng.uniform(-0.5, 0.5), rng.uniform(-0.3, 0.3), rng.uniform(1.0, 3.0)]).astype(np.float32)
    return out


def pose_error(T, T_ref):
    """(rotation angle in rad, translation distance in m) between two rigid transforms —
    MeasurePoseError semantics (include/geometry.h:147-180)."""
    dR = T[:3, :3] @ T_ref[:3, :3].T
    # atan2 of the skew part and the trace: accurate near zero (arccos alone bottoms out at ~1e-8)
    w = 0.5 * np.array([dR[2, 1] - dR[1, 2], dR[0, 2] - dR[2, 0], dR[1, 0] - dR[0, 1]])
    ang = np.arctan2(np.linalg.norm(w), (np.trace(dR) - 1.0) / 2.0)
    return float(ang), float(np.linalg.norm(T[:3, 3] - T_ref[:3, 3]))
