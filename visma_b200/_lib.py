"""ctypes loader for libvisma_b200.so (the C ABI declared in include/visma_b200.h).

There is no CPU fallback and no oracle import here: if the CUDA library is missing this module raises,
and every compute entry point reports VB200_ERR_NO_DEVICE on a box without a GPU.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# VISMA_B200_LIB: developer override to A/B an alternative build of the same CUDA library
LIB_PATH = os.environ.get("VISMA_B200_LIB") or os.path.join(_HERE, "libvisma_b200.so")

OK = 0
ERR_INVALID, ERR_NO_DEVICE, ERR_CUDA, ERR_NOMEM, ERR_NORMALS, ERR_DISTANCE = -1, -2, -3, -4, -5, -6
EST_P2P, EST_P2PLANE, EST_P2PLANE_GRAVITY = 0, 1, 2
OPT_NN_CACHE, OPT_SPLIT_TIMING = 1, 2

# every symbol include/visma_b200.h declares (tests/test_abi.py checks the two lists agree)
SYMBOLS = [
    "vb200_version", "vb200_strerror", "vb200_last_error", "vb200_device_count", "vb200_release_cached_memory",
    "vb200_scene_create", "vb200_scene_destroy", "vb200_scene_size", "vb200_scene_stream",
    "vb200_scene_sync", "vb200_knn1", "vb200_knn1_device", "vb200_knn1_bruteforce", "vb200_knn1_bruteforce_device", "vb200_icp_run", "vb200_batch_create",
    "vb200_batch_destroy", "vb200_batch_set_problems", "vb200_batch_run", "vb200_batch_results",
    "vb200_batch_corr", "vb200_batch_launches", "vb200_batch_iterate", "vb200_batch_set_option", "vb200_batch_last_kernel_ms", "vb200_batch_pass",
    "vb200_batch_set_totals_buffer", "vb200_batch_totals", "vb200_batch_solve", "vb200_estimate", "vb200_estimate_device", "vb200_rmse",
    "vb200_register_model_to_scene",
    "vb200_render_depth_batch", "vb200_render_depth_batch_ex", "vb200_render_edge_mask_batch", "vb200_voxel_downsample", "vb200_sample_mesh",
]


class VismaB200Error(RuntimeError):
    def __init__(self, status, where=""):
        self.status = status
        msg = lib().vb200_strerror(status).decode()
        if status == ERR_CUDA:
            msg += ": " + lib().vb200_last_error().decode()
        super().__init__("%s: %s (status %d)" % (where or "visma_b200", msg, status))


_lib = None


def build():
    """Compile the CUDA extension in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    import subprocess
    subprocess.check_call(["make", "-C", os.path.join(_HERE, "csrc"), "-j8"], stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "visma_b200: %s is missing — build it with `make -C visma_b200/csrc` "
            "(there is no CPU fallback)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    dp, ip, fp = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_float)
    i64p, vp = C.POINTER(C.c_int64), C.c_void_p
    L.vb200_version.restype = C.c_int
    L.vb200_strerror.restype = C.c_char_p
    L.vb200_strerror.argtypes = [C.c_int]
    L.vb200_last_error.restype = C.c_char_p
    L.vb200_device_count.restype = C.c_int
    L.vb200_release_cached_memory.argtypes = [C.c_int]
    L.vb200_scene_create.argtypes = [dp, dp, C.c_int64, C.c_double, C.c_int, C.POINTER(vp)]
    L.vb200_scene_destroy.argtypes = [vp]
    L.vb200_scene_size.argtypes = [vp, i64p, i64p, i64p, dp]
    L.vb200_scene_stream.restype = vp
    L.vb200_scene_stream.argtypes = [vp]
    L.vb200_scene_sync.argtypes = [vp]
    L.vb200_knn1.argtypes = [vp, dp, C.c_int64, C.c_double, ip, dp]
    L.vb200_knn1_device.argtypes = [vp, vp, C.c_int64, C.c_double, vp, vp]
    L.vb200_knn1_bruteforce.argtypes = [dp, C.c_int64, dp, C.c_int64, C.c_double, C.c_int, ip, dp]
    L.vb200_knn1_bruteforce_device.argtypes = [vp, C.c_int64, vp, C.c_int64, C.c_double, C.c_int, vp, vp, vp]
    L.vb200_icp_run.argtypes = [vp, dp, dp, i64p, C.c_int32, dp, C.c_int, dp, C.c_double, C.c_double,
                                C.c_double, C.c_int, dp, dp, dp, ip, ip, ip]
    L.vb200_batch_create.argtypes = [vp, dp, dp, i64p, C.c_int32, C.POINTER(vp)]
    L.vb200_batch_destroy.argtypes = [vp]
    L.vb200_batch_set_problems.argtypes = [vp, ip, dp, C.c_int32]
    L.vb200_batch_run.argtypes = [vp, C.c_int, dp, C.c_double, C.c_double, C.c_double, C.c_int]
    L.vb200_batch_results.argtypes = [vp, dp, dp, dp, ip, ip]
    L.vb200_batch_corr.argtypes = [vp, C.c_int32, ip, ip]
    L.vb200_batch_launches.restype = C.c_int64
    L.vb200_batch_launches.argtypes = [vp]
    L.vb200_batch_iterate.argtypes = [vp, C.c_int, dp, C.c_double, C.c_int]
    if hasattr(L, "vb200_batch_set_option"):  # (absent only in older builds loaded through VISMA_B200_LIB for A/B)
        L.vb200_batch_set_option.argtypes = [vp, C.c_int, C.c_int]
    L.vb200_batch_last_kernel_ms.argtypes = [vp, fp, fp]
    L.vb200_batch_pass.argtypes = [vp, C.c_int, C.c_double]
    L.vb200_batch_set_totals_buffer.argtypes = [vp, vp]
    L.vb200_batch_totals.restype = vp
    L.vb200_batch_totals.argtypes = [vp]
    L.vb200_batch_solve.argtypes = [vp, C.c_int, dp, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, i64p]
    L.vb200_estimate.argtypes = [dp, C.c_int64, dp, dp, C.c_int64, ip, C.c_int64, C.c_int, dp, C.c_int, dp]
    if hasattr(L, "vb200_rmse"):
        L.vb200_estimate_device.argtypes = [vp, C.c_int64, vp, vp, C.c_int64, vp, C.c_int64, C.c_int, dp, C.c_int, vp, dp]
        L.vb200_rmse.argtypes = [dp, C.c_int64, dp, C.c_int64, ip, C.c_int64, C.c_int, dp]
    L.vb200_register_model_to_scene.argtypes = [vp, dp, dp, C.c_int64, C.c_int, C.c_double, C.c_int, dp, ip, ip]
    L.vb200_render_depth_batch.argtypes = [fp, i64p, ip, i64p, C.c_int32, fp, fp, C.c_float, C.c_float,
                                           C.c_float, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int,
                                           C.c_int, C.POINTER(C.c_uint32), fp]
    L.vb200_render_depth_batch_ex.argtypes = [fp, i64p, ip, i64p, C.c_int32, fp, fp, C.c_float, C.c_float,
                                              C.c_float, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int,
                                              C.c_int, vp, vp, C.c_int, fp]
    L.vb200_render_edge_mask_batch.argtypes = [fp, i64p, ip, i64p, C.c_int32, fp, fp, C.c_float, C.c_float,
                                               C.c_float, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int,
                                               C.c_int, C.c_float, C.c_float, vp, vp, C.c_int]
    L.vb200_voxel_downsample.argtypes = [dp, dp, C.c_int64, C.c_double, C.c_int, dp, dp, i64p]
    L.vb200_sample_mesh.argtypes = [fp, C.c_int64, ip, C.c_int64, C.c_int64, C.c_uint64, C.c_int, dp, dp]
    _lib = L
    return L


def check(status, where=""):
    if status != OK:
        raise VismaB200Error(status, where)
