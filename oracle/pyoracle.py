"""ctypes bindings for oracle/libvisma_oracle.so — the plain-C restatement declared in oracle/oracle.h.

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
leg.  Never imported by visma_b200/.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvisma_oracle.so")

P2P, P2PLANE, P2PLANE_GRAVITY, P2P_CICP = 0, 1, 2, 3
ZMAX24 = (1 << 24) - 1

_lib = None


def build():
    subprocess.check_call(["make", "-C", _HERE, "oracle"], stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        _lib = C.CDLL(LIB_PATH)
        _lib.vo_index_create.restype = C.c_void_p
        _lib.vo_rmse_p2p.restype = C.c_double
        _lib.vo_linearize_depth.restype = C.c_float
        _lib.vo_voxel_downsample.restype = C.c_int64
    return _lib


def _p(a, t):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


def _d(a):
    return _p(a, C.c_double)


def _i(a):
    return _p(a, C.c_int32)


def _f(a):
    return _p(a, C.c_float)


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


class Index:
    """Exact uniform-grid NN index over a target cloud (stands in for KDTreeFlann)."""

    def __init__(self, tgt, cell):
        self.tgt = _f64(tgt)
        self.cell = float(cell)
        self.h = lib().vo_index_create(_d(self.tgt), C.c_int64(len(self.tgt)), C.c_double(cell))
        if not self.h:
            raise RuntimeError("vo_index_create failed")

    def close(self):
        if self.h:
            lib().vo_index_destroy(C.c_void_p(self.h))
            self.h = None

    def __del__(self):
        self.close()

    def knn1(self, q, radius):
        q = _f64(q)
        idx = np.empty(len(q), np.int32)
        d2 = np.empty(len(q), np.float64)
        rc = lib().vo_knn1(C.c_void_p(self.h), _d(q), C.c_int64(len(q)), C.c_double(radius),
                           _i(idx), _d(d2))
        if rc != 0:
            raise ValueError("vo_knn1 rc=%d" % rc)
        return idx, d2

    def registration_icp(self, src, max_dist, init=None, estimator=P2P, src_nrm=None, tgt_nrm=None,
                         gravity=(0.0, 1.0, 0.0), rel_fitness=1e-6, rel_rmse=1e-6, max_iter=30,
                         want_corr=False, want_trace=False):
        src, src_nrm, tgt_nrm = _f64(src), _f64(src_nrm), _f64(tgt_nrm)
        init = np.eye(4) if init is None else _f64(init)
        g = _f64(np.asarray(gravity))
        T = np.zeros((4, 4))
        fit, rmse, nc, its = C.c_double(), C.c_double(), C.c_int32(), C.c_int32()
        corr = np.zeros((max(len(src), 1), 2), np.int32) if want_corr else None
        trace = np.zeros((max_iter + 1, 19)) if want_trace else None
        rc = lib().vo_registration_icp(
            C.c_void_p(self.h), _d(self.tgt), _d(tgt_nrm), C.c_int64(len(self.tgt)), _d(src),
            _d(src_nrm), C.c_int64(len(src)), C.c_double(max_dist), _d(init), C.c_int(estimator),
            _d(g), C.c_double(rel_fitness), C.c_double(rel_rmse), C.c_int(max_iter), _d(T),
            C.byref(fit), C.byref(rmse), C.byref(nc), C.byref(its), _i(corr), _d(trace))
        out = dict(rc=rc, T=T, fitness=fit.value, rmse=rmse.value, ncorr=nc.value, iters=its.value)
        if want_corr:
            out["corr"] = corr[:nc.value]
        if want_trace:
            out["trace"] = trace[:its.value + 1]
        return out


def knn1_brute(tgt, q):
    tgt, q = _f64(tgt), _f64(q)
    idx = np.empty(len(q), np.int32)
    d2 = np.empty(len(q), np.float64)
    lib().vo_knn1_brute(_d(tgt), C.c_int64(len(tgt)), _d(q), C.c_int64(len(q)), _i(idx), _d(d2))
    return idx, d2


def estimate(src, tgt, corr, estimator=P2P, tgt_nrm=None, gravity=(0.0, 1.0, 0.0), with_scaling=False):
    src, tgt, tgt_nrm = _f64(src), _f64(tgt), _f64(tgt_nrm)
    corr = np.ascontiguousarray(corr, np.int32)
    T = np.zeros((4, 4))
    k = C.c_int64(len(corr))
    if estimator in (P2P, P2P_CICP):
        lib().vo_estimate_p2p(_d(src), _d(tgt), _i(corr), k, C.c_int(int(with_scaling)), _d(T))
    elif estimator == P2PLANE:
        lib().vo_estimate_p2plane(_d(src), _d(tgt), _d(tgt_nrm), _i(corr), k, _d(T))
    elif estimator == P2PLANE_GRAVITY:
        g = _f64(np.asarray(gravity))
        lib().vo_estimate_p2plane_gravity(_d(src), _d(tgt), _d(tgt_nrm), _i(corr), k, _d(g), _d(T))
    else:
        raise ValueError(estimator)
    return T


def rmse_p2p(src, tgt, corr):
    src, tgt = _f64(src), _f64(tgt)
    corr = np.ascontiguousarray(corr, np.int32)
    return lib().vo_rmse_p2p(_d(src), _d(tgt), _i(corr), C.c_int64(len(corr)))


def solve6(JTJ, JTr):
    JTJ, JTr = _f64(JTJ), _f64(JTr)
    x = np.zeros(6)
    ok = lib().vo_solve6(_d(JTJ), _d(JTr), _d(x))
    return bool(ok), x


def vec6_to_T(x):
    T = np.zeros((4, 4))
    lib().vo_vec6_to_T(_d(_f64(x)), _d(T))
    return T


def register_model_to_scene(model, scan, level=24, threshold=0.02, point_to_plane=False,
                            model_nrm=None, scan_nrm=None):
    model, scan, model_nrm, scan_nrm = _f64(model), _f64(scan), _f64(model_nrm), _f64(scan_nrm)
    T = np.zeros((4, 4))
    nc, best = C.c_int32(), C.c_int32()
    rc = lib().vo_register_model_to_scene(
        _d(scan), _d(scan_nrm), C.c_int64(len(scan)), _d(model), _d(model_nrm), C.c_int64(len(model)),
        C.c_int(level), C.c_double(threshold), C.c_int(int(point_to_plane)), _d(T), C.byref(nc),
        C.byref(best))
    assert rc == 0
    return dict(T=T, ncorr=nc.value, best_level=best.value)


# ---- renderer -------------------------------------------------------------------------------------
def projection(zn, zf, fx, fy, cx, cy, H, W):
    """Column-major float32[16] like glm (render/renderer.cpp:259-268)."""
    P = np.zeros(16, np.float32)
    lib().vo_projection(C.c_float(zn), C.c_float(zf), C.c_float(fx), C.c_float(fy), C.c_float(cx),
                        C.c_float(cy), C.c_int(H), C.c_int(W), _f(P))
    return P


def view(pose_colmajor):
    V = np.zeros(16, np.float32)
    lib().vo_view(_f(_f32(pose_colmajor).reshape(-1)), _f(V))
    return V


def render_depth(V, F, model_colmajor, view_colmajor, proj_colmajor, H, W):
    V = _f32(V)
    F = np.ascontiguousarray(F, np.int32)
    z24 = np.empty((H, W), np.uint32)
    depth = np.empty((H, W), np.float32)
    rc = lib().vo_render_depth(
        _f(V), C.c_int64(len(V)), _i(F), C.c_int64(len(F)), _f(_f32(model_colmajor).reshape(-1)),
        _f(_f32(view_colmajor).reshape(-1)), _f(_f32(proj_colmajor).reshape(-1)), C.c_int(H), C.c_int(W),
        _p(z24, C.c_uint32), _f(depth))
    assert rc == 0
    return z24, depth


def render_edge(z24, zn=0.05, zf=2.0):
    z24 = np.ascontiguousarray(z24, np.uint32)
    H, W = z24.shape
    out = np.empty((H, W), np.uint8)
    lib().vo_render_edge(_p(z24, C.c_uint32), C.c_int(H), C.c_int(W), C.c_float(zn), C.c_float(zf),
                         _p(out, C.c_uint8))
    return out


def render_mask(z24):
    z24 = np.ascontiguousarray(z24, np.uint32)
    H, W = z24.shape
    out = np.empty((H, W), np.uint8)
    lib().vo_render_mask(_p(z24, C.c_uint32), C.c_int(H), C.c_int(W), _p(out, C.c_uint8))
    return out


def linearize_depth(zb, zn, zf):
    return lib().vo_linearize_depth(C.c_float(zb), C.c_float(zn), C.c_float(zf))


def voxel_downsample(xyz, voxel, nrm=None):
    xyz, nrm = _f64(xyz), _f64(nrm)
    out = np.empty_like(xyz)
    out_n = np.empty_like(xyz) if nrm is not None else None
    k = lib().vo_voxel_downsample(_d(xyz), _d(nrm), C.c_int64(len(xyz)), C.c_double(voxel), _d(out),
                                  _d(out_n))
    if k < 0:
        raise ValueError("vo_voxel_downsample rc=%d" % k)
    return (out[:k], out_n[:k]) if nrm is not None else out[:k]


# ---- mesh surface sampling ---------------------------------------------------------------------------------
def _philox4x32_10(c0, c1, k0, k1):
    """Philox4x32-10 on numpy uint32 arrays (counter words c0, c1; c2 = c3 = 0), vectorised."""
    M0, M1, W0, W1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
    c0 = c0.astype(np.uint32); c1 = c1.astype(np.uint32)
    c2 = np.zeros_like(c0); c3 = np.zeros_like(c0)
    k0 = np.full_like(c0, k0); k1 = np.full_like(c0, k1)
    for _ in range(10):
        p0 = M0 * c0.astype(np.uint64)
        p1 = M1 * c2.astype(np.uint64)
        n0 = (p1 >> np.uint64(32)).astype(np.uint32) ^ c1 ^ k0
        n1 = (p1 & np.uint64(0xFFFFFFFF)).astype(np.uint32)
        n2 = (p0 >> np.uint64(32)).astype(np.uint32) ^ c3 ^ k1
        n3 = (p0 & np.uint64(0xFFFFFFFF)).astype(np.uint32)
        c0, c1, c2, c3 = n0, n1, n2, n3
        with np.errstate(over="ignore"):
            k0 = k0 + W0
            k1 = k1 + W1
    return c0, c1, c2, c3


def sample_mesh(V, F, n, seed=0):
    """Seeded restatement of feh::SamplePointCloudFromMesh (include/geometry.h:29-64) with the reference's two
    defects repaired (face index off by one, parallelogram instead of triangle) — the algorithm
    vb200_sample_mesh implements.  PARITY WITH THE REFERENCE IS STATISTICAL ONLY: it seeds from the wall clock
    (geometry.h:47).  Returns (points, unit face normals, face index)."""
    V = np.asarray(V, np.float32).astype(np.float64)
    F = np.asarray(F, np.int64)
    v0, e1, e2 = V[F[:, 0]], V[F[:, 1]] - V[F[:, 0]], V[F[:, 2]] - V[F[:, 0]]
    cr = np.cross(e1, e2)
    area = 0.5 * np.sqrt((cr[:, 0] * cr[:, 0] + cr[:, 1] * cr[:, 1]) + cr[:, 2] * cr[:, 2])
    cdf = np.cumsum(area)  # sequential double sum, as geometry.h:41-44
    i = np.arange(n, dtype=np.uint64)
    r0, r1, r2, r3 = _philox4x32_10((i & np.uint64(0xFFFFFFFF)).astype(np.uint32), (i >> np.uint64(32)).astype(np.uint32),
                                    np.uint32(seed & 0xFFFFFFFF), np.uint32((seed >> 32) & 0xFFFFFFFF))
    r = (r0.astype(np.float64) * 4294967296.0 + r1.astype(np.float64)) * (1.0 / 18446744073709551616.0) * cdf[-1]
    f = np.searchsorted(cdf, r, side="right").clip(0, len(F) - 1)
    a = r2.astype(np.float64) * (1.0 / 4294967296.0)
    b = r3.astype(np.float64) * (1.0 / 4294967296.0)
    fold = a + b > 1.0
    a = np.where(fold, 1.0 - a, a)
    b = np.where(fold, 1.0 - b, b)
    pts = (v0[f] + a[:, None] * e1[f]) + b[:, None] * e2[f]
    nr = cr[f] / np.maximum(np.linalg.norm(cr[f], axis=1, keepdims=True), 1e-300)
    return pts, nr, f
