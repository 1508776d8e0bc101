"""Python mirror of the reference's registration interface over the C ABI (tests and bench harness).

Names follow the reference: RegistrationICP / EvaluateRegistration / RegistrationResult /
ICPConvergenceCriteria (thirdparty/Open3D/src/Core/Registration/Registration.h:46-107),
TransformationEstimationPointToPoint / PointToPlane (TransformationEstimation.h:51-112),
cicp.TransformationEstimationPointToPoint4DoF (include/constrained_ICP.h:14-30), KDTreeFlann.SearchHybrid
(Core/Geometry/KDTreeFlann.h:73-75), RegisterModelToScene (src/annotation.cpp:29-64), VoxelDownSample
(Core/Geometry/DownSample.cpp:179-220).  The production host side for VISMA itself is the C++ adapter in
visma_b200/host/ (INTEGRATION.md); this module exists so the parity tests read like the reference's API.
"""
import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _lib
from ._lib import EST_P2P, EST_P2PLANE, EST_P2PLANE_GRAVITY, check, lib


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def _dp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_int32))


def _i64p(a):
    return a.ctypes.data_as(C.POINTER(C.c_int64))


@dataclass
class PointCloud:
    """open3d::PointCloud: points_/normals_ as N x 3 float64 (PointCloud.h:86-87)."""
    points_: np.ndarray
    normals_: np.ndarray = None

    def HasNormals(self):
        return self.normals_ is not None and len(self.normals_) == len(self.points_) and len(self.points_) > 0


@dataclass
class ICPConvergenceCriteria:
    relative_fitness_: float = 1e-6
    relative_rmse_: float = 1e-6
    max_iteration_: int = 30


@dataclass
class RegistrationResult:
    transformation_: np.ndarray = field(default_factory=lambda: np.eye(4))
    correspondence_set_: np.ndarray = field(default_factory=lambda: np.zeros((0, 2), np.int32))
    fitness_: float = 0.0
    inlier_rmse_: float = 0.0
    iterations_: int = 0  # not in the reference; number of estimator updates applied


class TransformationEstimationPointToPoint:
    kind = EST_P2P

    def __init__(self, with_scaling=False):
        if with_scaling:
            raise NotImplementedError("with_scaling is never enabled on VISMA's path")


class TransformationEstimationPointToPlane:
    kind = EST_P2PLANE


class TransformationEstimationPointToPoint4DoF(TransformationEstimationPointToPoint):
    """open3d::cicp::TransformationEstimationPointToPoint4DoF — in the reference a verbatim copy of the
    point-to-point estimator (src/constrained_ICP.cpp:25-37)."""


class TransformationEstimationPointToPlaneGravity:
    """The 4-DoF (yaw about gravity + translation) point-to-plane step the class name promised."""
    kind = EST_P2PLANE_GRAVITY

    def __init__(self, gravity_axis=(0.0, 1.0, 0.0)):
        self.gravity_axis = np.asarray(gravity_axis, np.float64)


class Scene:
    """A target cloud resident on one GPU with its NN grid — what KDTreeFlann::SetGeometry builds per
    RegistrationICP call in the reference (Registration.cpp:160-161), built once here."""

    def __init__(self, target, max_radius, device=0):
        pts = _f64(target.points_ if isinstance(target, PointCloud) else target)
        nrm = target.normals_ if isinstance(target, PointCloud) and target.HasNormals() else None
        nrm = _f64(nrm)
        self.n = len(pts)
        self.has_normals = nrm is not None
        self.max_radius = float(max_radius)
        self.device = device
        self._h = C.c_void_p()
        check(lib().vb200_scene_create(_dp(pts), _dp(nrm), len(pts), float(max_radius), device,
                                       C.byref(self._h)), "vb200_scene_create")

    def close(self):
        if self._h:
            lib().vb200_scene_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def size(self):
        n, nc, nf, cell = C.c_int64(), C.c_int64(), C.c_int64(), C.c_double()
        check(lib().vb200_scene_size(self._h, C.byref(n), C.byref(nc), C.byref(nf), C.byref(cell)))
        return dict(n=n.value, coarse_cells=nc.value, fine_cells=nf.value, cell=cell.value)

    def stream(self):
        return lib().vb200_scene_stream(self._h)

    def sync(self):
        check(lib().vb200_scene_sync(self._h), "vb200_scene_sync")

    # KDTreeFlann::SearchHybrid(q, radius, 1) for a batch of queries
    def SearchHybrid1(self, queries, radius):
        q = _f64(queries)
        idx = np.empty(len(q), np.int32)
        d2 = np.empty(len(q), np.float64)
        check(lib().vb200_knn1(self._h, _dp(q), len(q), float(radius), _ip(idx), _dp(d2)), "vb200_knn1")
        return idx, d2


def SearchHybrid1BruteForce(target, queries, radius, device=0):
    """KDTreeFlann::SearchHybrid(q, radius, 1) by exhaustive search (vb200_knn1_bruteforce): no index is
    built; every distance is taken in the reference's double arithmetic.  Same outputs as
    Scene.SearchHybrid1, bit for bit."""
    t, q = _f64(target), _f64(queries)
    idx = np.empty(len(q), np.int32)
    d2 = np.empty(len(q), np.float64)
    check(lib().vb200_knn1_bruteforce(_dp(t), len(t), _dp(q), len(q), float(radius), int(device), _ip(idx), _dp(d2)),
          "vb200_knn1_bruteforce")
    return idx, d2


class Batch:
    """Resident source clouds + ICP problems (cloud id, init) on a Scene."""

    def __init__(self, scene, sources, with_normals=None):
        self.scene = scene
        pts = [_f64(s.points_ if isinstance(s, PointCloud) else s) for s in sources]
        if with_normals is None:
            with_normals = len(sources) > 0 and all((isinstance(s, PointCloud) and s.HasNormals()) or len(p) == 0
                                                    for s, p in zip(sources, pts))
        self.sizes = [len(p) for p in pts]
        off = np.zeros(len(pts) + 1, np.int64)
        off[1:] = np.cumsum(self.sizes)
        self.offsets = off
        xyz = np.concatenate(pts) if pts else np.zeros((0, 3))
        xyz = _f64(xyz.reshape(-1, 3))
        # only the PRESENCE of source normals matters to the reference (Registration.cpp:152-157)
        nrm = xyz if with_normals else None
        self._h = C.c_void_p()
        check(lib().vb200_batch_create(scene.handle, _dp(xyz), _dp(nrm), _i64p(off), len(pts),
                                       C.byref(self._h)), "vb200_batch_create")
        self.P = 0
        self.cloud_ids = None

    def close(self):
        if self._h:
            lib().vb200_batch_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_problems(self, inits, cloud_ids=None):
        inits = _f64(np.asarray(inits).reshape(-1, 16))
        P = len(inits)
        ids = None if cloud_ids is None else np.ascontiguousarray(cloud_ids, np.int32)
        check(lib().vb200_batch_set_problems(self._h, _ip(ids), _dp(inits), P), "vb200_batch_set_problems")
        self.P = P
        self.cloud_ids = list(range(P)) if ids is None else [int(i) for i in ids]

    def run(self, estimation, max_dist, criteria=None):
        criteria = criteria or ICPConvergenceCriteria()
        g = getattr(estimation, "gravity_axis", None)
        g = _f64(g) if g is not None else None
        check(lib().vb200_batch_run(self._h, estimation.kind, _dp(g), float(max_dist),
                                    criteria.relative_fitness_, criteria.relative_rmse_,
                                    criteria.max_iteration_), "vb200_batch_run")

    def results(self, want_corr=False):
        P = self.P
        T = np.zeros((P, 4, 4))
        fit, rmse = np.zeros(P), np.zeros(P)
        nc, its = np.zeros(P, np.int32), np.zeros(P, np.int32)
        check(lib().vb200_batch_results(self._h, _dp(T), _dp(fit), _dp(rmse), _ip(nc), _ip(its)),
              "vb200_batch_results")
        out = []
        for p in range(P):
            r = RegistrationResult(T[p].copy(), np.zeros((0, 2), np.int32), float(fit[p]), float(rmse[p]),
                                   int(its[p]))
            if want_corr:
                m = self.sizes[self.cloud_ids[p]]
                corr = np.zeros((max(m, 1), 2), np.int32)
                k = C.c_int32()
                check(lib().vb200_batch_corr(self._h, p, _ip(corr), C.byref(k)), "vb200_batch_corr")
                r.correspondence_set_ = corr[:k.value].copy()
            else:
                r.correspondence_set_ = np.zeros((int(nc[p]), 0), np.int32)  # size only
            out.append(r)
        return out

    def launches(self):
        return int(lib().vb200_batch_launches(self._h))

    def iterate(self, estimation, max_dist, n_iter=1):
        """n unconditional ICP iterations from the current transforms (no convergence test); async."""
        g = getattr(estimation, "gravity_axis", None)
        g = _f64(g) if g is not None else None
        check(lib().vb200_batch_iterate(self._h, estimation.kind, _dp(g), float(max_dist), int(n_iter)),
              "vb200_batch_iterate")

    # --- iteration split at the cross-GPU exchange point (see include/visma_b200.h) ---
    def set_totals_buffer(self, device_ptr):
        check(lib().vb200_batch_set_totals_buffer(self._h, C.c_void_p(device_ptr)), "vb200_batch_set_totals_buffer")

    def pass_(self, estimation, max_dist):
        check(lib().vb200_batch_pass(self._h, estimation.kind, float(max_dist)), "vb200_batch_pass")

    def solve(self, estimation, max_dist, criteria, pass_index, npts_global=None):
        g = getattr(estimation, "gravity_axis", None)
        g = _f64(g) if g is not None else None
        n = None if npts_global is None else np.ascontiguousarray(npts_global, np.int64)
        check(lib().vb200_batch_solve(self._h, estimation.kind, _dp(g), float(max_dist), criteria.relative_fitness_,
                                      criteria.relative_rmse_, criteria.max_iteration_, int(pass_index),
                                      None if n is None else _i64p(n)), "vb200_batch_solve")

    def set_option(self, option, value):
        """measurement knobs: _lib.OPT_NN_CACHE (0 = search every point every pass), _lib.OPT_SPLIT_TIMING"""
        check(lib().vb200_batch_set_option(self._h, int(option), int(value)), "vb200_batch_set_option")

    def last_kernel_ms(self):
        a, b = C.c_float(), C.c_float()
        check(lib().vb200_batch_last_kernel_ms(self._h, C.byref(a), C.byref(b)), "vb200_batch_last_kernel_ms")
        return a.value, b.value


def RegistrationICPBatch(sources, scene, max_correspondence_distance, inits, estimation=None, criteria=None,
                         want_corr=True, packed=None):
    """open3d::RegistrationICP (Registration.h:102-107) for B independent sources through the single
    C-ABI call vb200_icp_run.  Mirrors the reference's error behaviour: on an invalid distance or missing
    normals every result is RegistrationResult(init)."""
    estimation = estimation or TransformationEstimationPointToPoint()
    criteria = criteria or ICPConvergenceCriteria()
    if packed is not None:
        # (xyz [sum M, 3] float64 C-contiguous — e.g. a pinned buffer —, offsets [B+1], has_normals): handed
        # to the C ABI as is, which is what a C++ caller with one contiguous allocation does
        xyz, off, has_n = packed
        off = np.ascontiguousarray(off, np.int64)
        B = len(off) - 1
        assert xyz.dtype == np.float64 and xyz.flags.c_contiguous
    else:
        B = len(sources)
        pts = [_f64(s.points_ if isinstance(s, PointCloud) else s) for s in sources]
        # an empty source cannot "have normals" in Open3D's sense and the reference would return init for it
        # either way; it must not switch the whole batch to the no-normals error path
        has_n = B > 0 and all((isinstance(s, PointCloud) and s.HasNormals()) or len(p) == 0
                              for s, p in zip(sources, pts))
        off = np.zeros(B + 1, np.int64)
        off[1:] = np.cumsum([len(p) for p in pts])
        xyz = _f64(np.concatenate(pts).reshape(-1, 3)) if B else np.zeros((0, 3))
    inits = _f64(np.asarray(inits).reshape(-1, 16))
    T = np.zeros((B, 4, 4))
    fit, rmse = np.zeros(B), np.zeros(B)
    nc, its = np.zeros(B, np.int32), np.zeros(B, np.int32)
    corr = np.zeros((max(int(off[-1]), 1), 2), np.int32) if want_corr else None
    g = getattr(estimation, "gravity_axis", None)
    g = _f64(g) if g is not None else None
    rc = lib().vb200_icp_run(scene.handle, _dp(xyz), _dp(xyz if has_n else None), _i64p(off), B, _dp(inits),
                             estimation.kind, _dp(g), float(max_correspondence_distance),
                             criteria.relative_fitness_, criteria.relative_rmse_, criteria.max_iteration_,
                             _dp(T), _dp(fit), _dp(rmse), _ip(nc), _ip(its), _ip(corr))
    if rc not in (_lib.OK, _lib.ERR_DISTANCE, _lib.ERR_NORMALS):
        check(rc, "vb200_icp_run")
    out = []
    for b in range(B):
        cs = corr[off[b]:off[b] + nc[b]].copy() if want_corr else np.zeros((int(nc[b]), 0), np.int32)
        out.append(RegistrationResult(T[b].copy(), cs, float(fit[b]), float(rmse[b]), int(its[b])))
    return out


def RegistrationICP(source, target_scene, max_correspondence_distance, init=None, estimation=None,
                    criteria=None):
    """Single-source form with the reference's argument order (target is a resident Scene)."""
    init = np.eye(4) if init is None else init
    return RegistrationICPBatch([source], target_scene, max_correspondence_distance, [init], estimation,
                                criteria)[0]


def EvaluateRegistration(source, target_scene, max_correspondence_distance, transformation=None):
    """open3d::EvaluateRegistration (Registration.cpp:127-139): one correspondence pass, no update."""
    T = np.eye(4) if transformation is None else transformation
    return RegistrationICP(source, target_scene, max_correspondence_distance, T,
                           TransformationEstimationPointToPoint(), ICPConvergenceCriteria(1e-6, 1e-6, 0))


def ComputeTransformation(estimation, source, target, corres, device=0):
    """TransformationEstimation::ComputeTransformation(source, target, corres) on the GPU."""
    s = _f64(source.points_ if isinstance(source, PointCloud) else source)
    t = _f64(target.points_ if isinstance(target, PointCloud) else target)
    tn = _f64(target.normals_) if isinstance(target, PointCloud) and target.HasNormals() else None
    corr = np.ascontiguousarray(corres, np.int32).reshape(-1, 2)
    g = getattr(estimation, "gravity_axis", None)
    g = _f64(g) if g is not None else None
    T = np.zeros((4, 4))
    check(lib().vb200_estimate(_dp(s), len(s), _dp(t), _dp(tn), len(t), _ip(corr), len(corr),
                               estimation.kind, _dp(g), device, _dp(T)), "vb200_estimate")
    return T


def ComputeRMSE(source, target, corres, device=0):
    """cicp::TransformationEstimationPointToPoint4DoF::ComputeRMSE (src/constrained_ICP.cpp:13-23) ==
    TransformationEstimationPointToPoint::ComputeRMSE: sqrt(sum |s - t|^2 / K) on the GPU; 0 for an empty set."""
    s = _f64(source.points_ if isinstance(source, PointCloud) else source)
    t = _f64(target.points_ if isinstance(target, PointCloud) else target)
    corr = np.ascontiguousarray(corres, np.int32).reshape(-1, 2)
    out = C.c_double()
    check(lib().vb200_rmse(_dp(s), len(s), _dp(t), len(t), _ip(corr), len(corr), device, C.byref(out)), "vb200_rmse")
    return out.value


def ComputeTransformationDevice(estimation, d_src, m, d_tgt, d_tgt_nrm, n, d_corr, K, device=0, stream=None):
    """ComputeTransformation with every array already on the GPU (raw device pointers as ints): the rows are
    gathered by the kernel (vb200_estimate_device)."""
    g = getattr(estimation, "gravity_axis", None)
    g = _f64(g) if g is not None else None
    T = np.zeros((4, 4))
    check(lib().vb200_estimate_device(C.c_void_p(d_src), int(m), C.c_void_p(d_tgt), C.c_void_p(d_tgt_nrm or None), int(n),
                                      C.c_void_p(d_corr), int(K), estimation.kind, _dp(g), device,
                                      C.c_void_p(stream or None), _dp(T)), "vb200_estimate_device")
    return T


def RegisterModelToScene(model, scan_scene, rotation_level=24, distance_threshold=0.02, point_to_plane=False):
    """feh::RegisterModelToScene (src/annotation.cpp:29-64) with all yaw inits batched on the GPU."""
    m = _f64(model.points_ if isinstance(model, PointCloud) else model)
    has_n = isinstance(model, PointCloud) and model.HasNormals()
    T = np.zeros((4, 4))
    nc, best = C.c_int32(), C.c_int32()
    rc = lib().vb200_register_model_to_scene(scan_scene.handle, _dp(m), _dp(m if has_n else None), len(m),
                                             int(rotation_level), float(distance_threshold),
                                             1 if point_to_plane else 0, _dp(T), C.byref(nc), C.byref(best))
    if rc not in (_lib.OK, _lib.ERR_DISTANCE, _lib.ERR_NORMALS):
        check(rc, "vb200_register_model_to_scene")
    return dict(T=T, ncorr=nc.value, best_level=best.value, status=rc)


def VoxelDownSample(cloud, voxel_size, device=0):
    """open3d::VoxelDownSample (DownSample.cpp:179-220); output ordered by voxel index."""
    pts = _f64(cloud.points_ if isinstance(cloud, PointCloud) else cloud)
    nrm = _f64(cloud.normals_) if isinstance(cloud, PointCloud) and cloud.HasNormals() else None
    out = np.empty_like(pts)
    out_n = np.empty_like(pts) if nrm is not None else None
    k = C.c_int64()
    rc = lib().vb200_voxel_downsample(_dp(pts), _dp(nrm), len(pts), float(voxel_size), device, _dp(out),
                                      _dp(out_n), C.byref(k))
    if rc == _lib.ERR_INVALID and not voxel_size > 0:
        return PointCloud(np.zeros((0, 3)))  # reference returns an empty cloud
    check(rc, "vb200_voxel_downsample")
    return PointCloud(out[:k.value].copy(), None if out_n is None else out_n[:k.value].copy())


def SamplePointCloudFromMesh(V, F, max_num_pts=1000, seed=0, device=0, with_normals=False):
    """feh::SamplePointCloudFromMesh (include/geometry.h:29-64) on the GPU: area-weighted surface samples,
    reproducible from `seed`.  Returns points (and unit face normals when asked)."""
    V = np.ascontiguousarray(np.asarray(V, np.float32).reshape(-1, 3))
    F = np.ascontiguousarray(np.asarray(F, np.int32).reshape(-1, 3))
    out = np.empty((int(max_num_pts), 3), np.float64)
    nrm = np.empty((int(max_num_pts), 3), np.float64) if with_normals else None
    check(lib().vb200_sample_mesh(V.ctypes.data_as(C.POINTER(C.c_float)), len(V), _ip(F), len(F), int(max_num_pts),
                                  int(seed), device, _dp(out), _dp(nrm)), "vb200_sample_mesh")
    return (out, nrm) if with_normals else out
