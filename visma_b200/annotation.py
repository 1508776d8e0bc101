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
