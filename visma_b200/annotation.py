"""Host-side mirror of VISMA's orientation-constrained annotation driver (src/annotation.cpp:66-176) on top
of the GPU operators: gravity alignment from the floor fragment, per-object centring, the 24-yaw
RegisterModelToScene search, and the composition of the final model -> scene transform.  File IO (PLY / OBJ /
JSON through Open3D, libigl, folly) is out of scope: callers hand in arrays and get a dict of 3x4 matrices in
the layout WriteMatrixToJson stores (core/utils.h:333-339).

Only small host math lives here (3x3 SVD, 4x4 products); the heavy steps are vb200_voxel_downsample and
vb200_register_model_to_scene.
"""
import numpy as np

from . import registration as reg


def FindPlaneNormal(pts):
    """include/geometry.h:18-26: smallest singular vector of the points' 3x3 covariance.  The reference takes
    whatever sign JacobiSVD returns; here the normal is oriented towards +Y so the gravity rotation below
    never turns the scene upside down."""
    pts = np.asarray(pts, np.float64)
    c = pts - pts.mean(0)
    P = c.T @ c / len(pts)
    _, _, Vt = np.linalg.svd(P)
    n = Vt[2] / np.linalg.norm(Vt[2])
    return n if n[1] >= 0 else -n


def RotationBetweenVectors(u, v):
    """core/utils.h:229-233: Eigen::Quaternion::FromTwoVectors(u, v).toRotationMatrix()."""
    u = np.asarray(u, np.float64) / np.linalg.norm(u)
    v = np.asarray(v, np.float64) / np.linalg.norm(v)
    c = float(u @ v)
    if c < -1 + 1e-12:  # opposite vectors: rotate by pi about any axis orthogonal to u
        a = np.cross(u, [1.0, 0, 0] if abs(u[0]) < 0.9 else [0, 1.0, 0])
        a /= np.linalg.norm(a)
        return 2 * np.outer(a, a) - np.eye(3)
    w = np.cross(u, v)
    K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    return np.eye(3) + K + K @ K / (1 + c)


def MinY(points):
    """src/annotation.cpp:21-27"""
    return float(np.asarray(points)[:, 1].min())


def _T(R=None, t=None):
    T = np.eye(4)
    if R is not None:
        T[:3, :3] = R
    if t is not None:
        T[:3, 3] = t
    return T


def _apply(T, pts):
    return pts @ T[:3, :3].T + T[:3, 3]


def GravityAlignment(floor_points):
    """src/annotation.cpp:80-89: T0 rotates the floor normal onto +Y."""
    return _T(RotationBetweenVectors(FindPlaneNormal(floor_points), [0.0, 1.0, 0.0]))


def AnnotateObject(scan_points, model_points, T0, icp, device=0, register=None):
    """One iteration of the loop at src/annotation.cpp:103-153.
    scan_points: the object's fragment of the scene (N x 3); model_points: CAD surface samples (the reference
    draws 2*|scan| of them, :126); icp: the "ICP" block of cfg/tool.json (voxel_size, point_to_plane,
    rotation_level, distance_threshold).  Returns (Ttot 4x4 model -> scene, details).
    `register` lets tests substitute another RegisterModelToScene (the CPU oracle) for the GPU one."""
    scan = reg.VoxelDownSample(np.asarray(scan_points, np.float64), icp["voxel_size"], device).points_  # :112
    scan = _apply(T0, scan)                                                                      # :113
    mean = scan.mean(0)
    T1 = _T(t=[-mean[0], -MinY(scan), -mean[2]])                                                 # :115-119
    scan = _apply(T1, scan)
    model = np.asarray(model_points, np.float64)
    mean = model.mean(0)
    T2 = _T(t=[-mean[0], -MinY(model), -mean[2]])                                                # :128-132
    model = _apply(T2, model)
    if register is None:
        scene = reg.Scene(scan, icp["distance_threshold"], device)
        out = reg.RegisterModelToScene(model, scene, icp["rotation_level"], icp["distance_threshold"],
                                       bool(icp.get("point_to_plane", False)))
        scene.close()
    else:
        out = register(model, scan)
    T3 = out["T"]                                                                                # :141
    T10 = T1 @ T0
    inv = np.eye(4)                                                                              # :147-149
    inv[:3, :3] = T10[:3, :3].T
    inv[:3, 3] = -T10[:3, :3].T @ T10[:3, 3]
    Ttot = inv @ T3 @ T2                                                                         # :150
    return Ttot, dict(T1=T1, T2=T2, T3=T3, ncorr=out["ncorr"], best_level=out["best_level"], n_scan=len(scan))


def AnnotationTool(floor_points, scans, models, config, device=0, register=None):
    """src/annotation.cpp:66-176 without the file IO.  scans: {scan_name: N x 3}; models: {model_name: M x 3
    surface samples}, model_name = scan_name up to its last '_' (:108-109).  Returns {scan_name: 3 x 4}."""
    T0 = GravityAlignment(floor_points)
    out = {}
    for scan_name, pts in scans.items():
        model_name = scan_name[:scan_name.rfind("_")]
        Ttot, _ = AnnotateObject(pts, models[model_name], T0, config["ICP"], device, register)
        out[scan_name] = Ttot[:3, :4].copy()
    return out


def AnnotationToolFromFiles(config, device=0, register=None, seed=0):
    """feh::AnnotationTool(config) with its file IO (src/annotation.cpp:66-176): reads <dataroot>/<dataset>/
    fragments/{floor.ply, objects.json, <entry>.ply} and <CAD_database_root>/<model>.obj, samples 2 x |scan| model
    points per object (:126; seeded here, the reference seeds from the clock), runs the orientation-constrained
    registration and writes fragments/alignment.json (key -> 3x4 row-major, :153,175).  Returns the dict."""
    import os
    from . import io3d
    scene_dir = os.path.join(config["dataroot"], config["dataset"])
    frag = os.path.join(scene_dir, "fragments")
    floor, _ = io3d.read_ply(os.path.join(frag, "floor.ply"))
    T0 = GravityAlignment(floor)
    out = {}
    for k, scan_name in enumerate(io3d.load_json(os.path.join(frag, "objects.json"))["entries"]):
        model_name = scan_name[:scan_name.rfind("_")]
        scan, _ = io3d.read_ply(os.path.join(frag, scan_name + ".ply"))
        V, F = io3d.read_obj(os.path.join(config["CAD_database_root"], model_name + ".obj"))
        n_scan = len(reg.VoxelDownSample(scan, config["ICP"]["voxel_size"], device).points_)
        model = reg.SamplePointCloudFromMesh(V, F, 2 * n_scan, seed=seed + k, device=device)
        Ttot, _ = AnnotateObject(scan, model, T0, config["ICP"], device, register)
        out[scan_name] = Ttot[:3, :4].copy()
    io3d.save_json({k: io3d.matrix_to_json(v) for k, v in out.items()}, os.path.join(frag, "alignment.json"))
    return out


def ComputeErrorMetric(errors):
    """include/geometry.h:85-101 (median = element n >> 1 of the sorted errors)."""
    e = np.asarray(errors, np.float64)
    if len(e) == 0:
        return dict(mean=float("nan"), std=float("nan"), median=float("nan"), min=float("inf"), max=float("-inf"))
    mean = float(e.sum() / len(e))
    return dict(mean=mean, std=float(np.sqrt(max((e * e).sum() / len(e) - mean * mean, 0.0))),
                median=float(np.sort(e)[len(e) >> 1]), min=float(e.min()), max=float(e.max()))


def MeasurePoseError(Gs, Gt, dist_thresh=0.5):
    """include/geometry.h:147-180, as written: for every source pose the target poses are scanned for the nearest
    translation below dist_thresh, and an error pair (translation distance, rotation angle of Rt^T Rs) is collected
    INSIDE the scan — once per remaining target after the first match, each time against the best so far — so a
    source matched early is counted several times.  Reproduced, since the reference's statistics include it.
    Gs, Gt: lists of 3x4.  Returns (translation metric, rotation metric [rad])."""
    t_err, r_err = [], []
    for S in Gs:
        best, idx = dist_thresh, -1
        for j, T in enumerate(Gt):
            dn = float(np.linalg.norm(T[:3, 3] - S[:3, 3]))
            if dn < best:
                best, idx = dn, j
            if idx != -1:
                dR = Gt[idx][:3, :3].T @ S[:3, :3]
                w = 0.5 * np.array([dR[2, 1] - dR[1, 2], dR[0, 2] - dR[2, 0], dR[1, 0] - dR[0, 1]])
                t_err.append(float(np.linalg.norm(Gt[idx][:3, 3] - S[:3, 3])))
                r_err.append(float(np.arctan2(np.linalg.norm(w), (np.trace(dR) - 1.0) / 2.0)))  # AngleAxis(dR).angle()
    return ComputeErrorMetric(t_err), ComputeErrorMetric(r_err)


def _se3_log(T):
    """6-vector (rho, phi) with exp(.) = T: SO(3) log by core/rodrigues.h:184-226 (invrodrigues), translation by V^-1."""
    R, t = T[:3, :3], T[:3, 3]
    c = 0.5 * (np.trace(R) - 1.0)
    vee = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    if c > 1.0 - 1e-10:
        w = 0.5 * vee
    else:
        th = np.arccos(max(c, -1.0))
        w = th * 0.5 * vee / np.sin(th)
    th = np.linalg.norm(w)
    K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    if th < 1e-8:
        Vinv = np.eye(3) - 0.5 * K + K @ K / 12.0
    else:
        Vinv = np.eye(3) - 0.5 * K + (1.0 / th ** 2 - (1.0 + np.cos(th)) / (2.0 * th * np.sin(th))) * K @ K
    return np.concatenate([Vinv @ t, w])


def _se3_exp(x):
    rho, w = np.asarray(x[:3], np.float64), np.asarray(x[3:], np.float64)
    th = np.linalg.norm(w)
    K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    if th < 1e-8:
        R = np.eye(3) + K + 0.5 * K @ K
        V = np.eye(3) + 0.5 * K + K @ K / 6.0
    else:
        R = np.eye(3) + np.sin(th) / th * K + (1 - np.cos(th)) / th ** 2 * K @ K
        V = np.eye(3) + (1 - np.cos(th)) / th ** 2 * K + (th - np.sin(th)) / th ** 3 * K @ K
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = V @ rho
    return T


def OptimizeAlignment(tgt, src, matches, max_iter=100):
    """feh::OptimizeAlignment (src/evaluation.cpp:43-77): the reference `throw`s here and keeps the intended
    algorithm in a comment — an iteratively re-weighted mean of the relative poses dT_k = tgt_k * src_k^-1 in the
    tangent space of SE(3): sum = sum_k w_k log(dT_k); T = exp(sum); w_k ~ 1 / max(1e-4, |log(tgt_k (T src_k)^-1)|),
    normalised; stop when the sum changes by < 1e-5 relative; at most 100 iterations.  Implemented as written (in
    double; the comment casts to float).  tgt / src: {id: 4x4 model_to_scene}; matches: [(src id, tgt id)]."""
    if len(matches) == 0:
        return np.eye(4)
    w = np.full(len(matches), 1.0 / len(matches))
    last = None
    total = np.zeros(6)
    for _ in range(max_iter):
        total = np.zeros(6)
        for k, (i, j) in enumerate(matches):
            total += w[k] * _se3_log(tgt[j] @ np.linalg.inv(src[i]))
        T = _se3_exp(total)
        for k, (i, j) in enumerate(matches):
            w[k] = 1.0 / max(1e-4, np.linalg.norm(_se3_log(tgt[j] @ np.linalg.inv(T @ src[i]))))
        w /= w.sum()
        if last is not None and np.linalg.norm(last - total) / max(np.linalg.norm(total), 1e-300) < 1e-5:
            break
        last = total
    return _se3_exp(total)


def FindCorrespondence(tgt, src, T_tgt_src, threshold):
    """src/evaluation.cpp:18-41.  tgt / src: {id: (model_name, 4x4 model_to_scene)}; -> [(src id, tgt id)]."""
    matches = []
    for i, (_, Ts) in src.items():
        best, best_j = threshold, -1
        for j, (_, Tt) in tgt.items():
            dT = np.linalg.inv(T_tgt_src @ Ts) @ Tt
            n = float(np.linalg.norm(dT[:3, 3]))
            if n < best:
                best, best_j = n, j
        if best_j >= 0:
            matches.append((i, best_j))
    return matches


def RegisterScenes(tgt, src):
    """src/evaluation.cpp:80-112 with the working OptimizeAlignment: -> (T_tgt_src 4x4, matches)."""
    best = []
    for i, (name_s, Ts) in src.items():
        for j, (name_t, Tt) in tgt.items():
            if name_s == name_t:
                m = FindCorrespondence(tgt, src, Tt @ np.linalg.inv(Ts), 0.5)
                if len(m) > len(best):
                    best = m
    T = OptimizeAlignment({j: T for j, (_, T) in tgt.items()}, {i: T for i, (_, T) in src.items()}, best)
    return T, best


def ICPRefinement(scene_points, model_clouds, model_to_scene, T_scene_src, options, device=0):
    """feh::ICPRefinement (src/evaluation.cpp:244-274): scene_est = union of the posed model samples
    (:250-256), scene voxel-down-sampled (:258), ONE RegistrationICP(scene_est -> scene) from T_scene_src
    (:260-271).  model_clouds: list of M_i x 3 samples in model frames; model_to_scene: list of 4x4.
    The point-to-plane branch needs normals on both clouds; scene_est has none in the reference, which therefore
    returns T_scene_src unchanged (Registration.cpp:152-157) — reproduced."""
    est = np.concatenate([_apply(np.asarray(T), np.asarray(p, np.float64))
                          for p, T in zip(model_clouds, model_to_scene)])
    scene = reg.VoxelDownSample(np.asarray(scene_points, np.float64), options.get("voxel_size", 0.02), device)
    sc = reg.Scene(scene, options.get("max_distance", 0.05), device)
    estimation = (reg.TransformationEstimationPointToPlane() if options.get("use_point_to_plane")
                  else reg.TransformationEstimationPointToPoint())
    res = reg.RegistrationICP(reg.PointCloud(est), sc, options.get("max_distance", 0.05), T_scene_src, estimation)
    sc.close()
    return res, scene
