"""ctypes bindings for oracle/_ref/libvisma_ref.so — the UNMODIFIED reference
(Open3D 0.3.0 + FLANN + Eigen + VISMA constrained_ICP.cpp) behind oracle/ref_shim.cpp.

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  Never imported by visma_b200/.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libvisma_ref.so")

P2P, P2PLANE, CICP_4DOF = 0, 1, 2

_lib = None


def available():
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(LIB_PATH)
        _lib.ref_kdtree_create.restype = C.c_void_p
        _lib.ref_voxel_downsample.restype = C.c_int64
    return _lib


def _d(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _i(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_int32))


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def num_threads():
    return lib().ref_num_threads()


def set_num_threads(n):
    lib().ref_set_num_threads(int(n))


def registration_icp(src, tgt, max_dist, init=None, estimator=P2P, src_nrm=None,
                     tgt_nrm=None, rel_fitness=1e-6, rel_rmse=1e-6, max_iter=30,
                     want_corr=False):
    """open3d::RegistrationICP (Registration.cpp:141-186)."""
    src, tgt, src_nrm, tgt_nrm = _f64(src), _f64(tgt), _f64(src_nrm), _f64(tgt_nrm)
    init = np.eye(4) if init is None else _f64(init)
    T = np.zeros((4, 4))
    fit, rmse, nc = C.c_double(), C.c_double(), C.c_int32()
    corr = np.zeros((len(src), 2), np.int32) if want_corr else None
    rc = lib().ref_registration_icp(
        _d(src), _d(src_nrm), C.c_int64(len(src)), _d(tgt), _d(tgt_nrm),
        C.c_int64(len(tgt)), C.c_double(max_dist), _d(init), C.c_int(estimator),
        C.c_double(rel_fitness), C.c_double(rel_rmse), C.c_int(max_iter), _d(T),
        C.byref(fit), C.byref(rmse), C.byref(nc), _i(corr))
    assert rc == 0
    out = dict(T=T, fitness=fit.value, rmse=rmse.value, ncorr=nc.value)
    if want_corr:
        c = corr[:nc.value]
        out["corr"] = c[np.argsort(c[:, 0], kind="stable")]
    return out


def evaluate_registration(src, tgt, max_dist, T=None, want_corr=False):
    """open3d::EvaluateRegistration (Registration.cpp:127-139)."""
    src, tgt = _f64(src), _f64(tgt)
    T = np.eye(4) if T is None else _f64(T)
    fit, rmse, nc = C.c_double(), C.c_double(), C.c_int32()
    corr = np.zeros((len(src), 2), np.int32) if want_corr else None
    rc = lib().ref_evaluate_registration(
        _d(src), C.c_int64(len(src)), _d(tgt), C.c_int64(len(tgt)),
        C.c_double(max_dist), _d(T), C.byref(fit), C.byref(rmse), C.byref(nc), _i(corr))
    assert rc == 0
    out = dict(fitness=fit.value, rmse=rmse.value, ncorr=nc.value)
    if want_corr:
        c = corr[:nc.value]
        out["corr"] = c[np.argsort(c[:, 0], kind="stable")]
    return out


class KDTree:
    """open3d::KDTreeFlann over a target cloud (KDTreeFlann.cpp:191-208)."""

    def __init__(self, tgt):
        self.tgt = _f64(tgt)
        self.h = lib().ref_kdtree_create(_d(self.tgt), C.c_int64(len(self.tgt)))
        if not self.h:
            raise RuntimeError("KDTreeFlann::SetGeometry failed")

    def close(self):
        if self.h:
            lib().ref_kdtree_destroy(C.c_void_p(self.h))
            self.h = None

    def __del__(self):
        self.close()

    def search_hybrid1(self, q, radius):
        """SearchHybrid(q, radius, 1) per query; idx -1 where none."""
        q = _f64(q)
        idx = np.empty(len(q), np.int32)
        d2 = np.empty(len(q), np.float64)
        rc = lib().ref_kdtree_search_hybrid1(C.c_void_p(self.h), _d(q), C.c_int64(len(q)),
                                             C.c_double(radius), _i(idx), _d(d2))
        assert rc == 0
        return idx, d2

    def search_knn1(self, q):
        q = _f64(q)
        idx = np.empty(len(q), np.int32)
        d2 = np.empty(len(q), np.float64)
        rc = lib().ref_kdtree_search_knn1(C.c_void_p(self.h), _d(q), C.c_int64(len(q)),
                                          _i(idx), _d(d2))
        assert rc == 0
        return idx, d2

    def icp_trace(self, src, max_dist, init=None, estimator=P2P, src_nrm=None, tgt_nrm=None,
                  rel_fitness=1e-6, rel_rmse=1e-6, max_iter=30):
        """Registration.cpp:159-185 re-driven so each iteration is observable.
        Returns rows [fitness, rmse, ncorr, T(16)] for iterations 0..iters."""
        src, src_nrm, tgt_nrm = _f64(src), _f64(src_nrm), _f64(tgt_nrm)
        init = np.eye(4) if init is None else _f64(init)
        trace = np.zeros((max_iter + 1, 19))
        it = lib().ref_icp_trace(
            C.c_void_p(self.h), _d(src), _d(src_nrm), C.c_int64(len(src)), _d(tgt_nrm),
            C.c_double(max_dist), _d(init), C.c_int(estimator), C.c_double(rel_fitness),
            C.c_double(rel_rmse), C.c_int(max_iter), _d(trace))
        assert it >= 0
        return trace[:it + 1]


def estimate(src, tgt, corr, estimator=P2P, tgt_nrm=None):
    """TransformationEstimation::ComputeTransformation (TransformationEstimation.cpp:47-103,
    src/constrained_ICP.cpp:25-37)."""
    src, tgt, tgt_nrm = _f64(src), _f64(tgt), _f64(tgt_nrm)
    corr = np.ascontiguousarray(corr, np.int32)
    T = np.zeros((4, 4))
    rc = lib().ref_estimate(_d(src), C.c_int64(len(src)), _d(tgt), _d(tgt_nrm),
                            C.c_int64(len(tgt)), _i(corr), C.c_int64(len(corr)),
                            C.c_int(estimator), _d(T))
    assert rc == 0
    return T


def voxel_downsample(xyz, voxel, nrm=None):
    """open3d::VoxelDownSample (DownSample.cpp:179-220)."""
    xyz, nrm = _f64(xyz), _f64(nrm)
    out = np.empty_like(xyz)
    out_n = np.empty_like(xyz) if nrm is not None else None
    k = lib().ref_voxel_downsample(_d(xyz), _d(nrm), C.c_int64(len(xyz)), C.c_double(voxel),
                                   _d(out), _d(out_n))
    return (out[:k], out_n[:k]) if nrm is not None else out[:k]


def register_model_to_scene(model, scan, level=24, threshold=0.02, point_to_plane=False,
                            model_nrm=None, scan_nrm=None):
    """feh::RegisterModelToScene (src/annotation.cpp:29-64) around the real RegistrationICP."""
    model, scan, model_nrm, scan_nrm = _f64(model), _f64(scan), _f64(model_nrm), _f64(scan_nrm)
    T = np.zeros((4, 4))
    nc, best = C.c_int32(), C.c_int32()
    rc = lib().ref_register_model_to_scene(
        _d(model), _d(model_nrm), C.c_int64(len(model)), _d(scan), _d(scan_nrm),
        C.c_int64(len(scan)), C.c_int(level), C.c_double(threshold),
        C.c_int(1 if point_to_plane else 0), _d(T), C.byref(nc), C.byref(best))
    assert rc == 0
    return dict(T=T, ncorr=nc.value, best_level=best.value)


def transform(xyz, T, nrm=None):
    xyz = _f64(xyz).copy()
    nrm = None if nrm is None else _f64(nrm).copy()
    lib().ref_transform(_d(xyz), _d(nrm), C.c_int64(len(xyz)), _d(_f64(T)))
    return xyz if nrm is None else (xyz, nrm)
