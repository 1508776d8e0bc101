"""dev tool: per-iteration search-path counters of a -DVB_STATS build on the BASELINE workload.
usage: VISMA_B200_LIB=build/variants/lib_stats.so python scripts/search_stats.py [n_iter]"""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from visma_b200 import _lib, registration as reg, synth
n_iter = int(sys.argv[1]) if len(sys.argv) > 1 else 12
d = synth.make_room_scene(2_000_000, 32, 50_000, source_seed=0)
scene = reg.Scene(reg.PointCloud(d["scene_xyz"], d["scene_nrm"]), 0.075, device=0)
batch = reg.Batch(scene, [reg.PointCloud(p, n) for p, n in d["sources"]])
batch.set_problems(d["T_init"])
L = _lib.lib()
L.vb200_debug_stats.argtypes = [C.POINTER(C.c_ulonglong), C.c_int]
buf = (C.c_ulonglong * 24)()
L.vb200_debug_states.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float)]
cum, lm = (C.c_float * 32)(), (C.c_float * 32)()
L.vb200_debug_stats(None, 1)
names = ["valid", "prior", "", "coop", "coop:reach", "coop:overflow", "coop:nohome", "runs", "lane_steps", "warp_max_steps",
         "warps_w_coop", "warps", "A:hard", "A:settled", "A:set_tried", "A:set_ok", "B:top5", "B:set_made", "B:set_by_e4",
         "A:has_set", "B:eligible", "B:near_overflow", "B:near_lt2", "B:not_single_round"]
est = reg.TransformationEstimationPointToPlane()
for it in range(n_iter):
    batch.iterate(est, 0.075, 1)
    scene.sync()
    L.vb200_debug_stats(buf, 1)
    v = list(buf)
    L.vb200_debug_states(batch._h, cum, lm)
    lmv = np.array(list(lm))
    print("   last_move mm: min %.3f med %.3f max %.3f | cum med %.3f" % (lmv.min() * 1e3, np.median(lmv) * 1e3, lmv.max() * 1e3, np.median(list(cum)) * 1e3))
    print(it, " ".join("%s=%d" % (n, v[i]) for i, n in enumerate(names) if n),
          "| steps/lane %.1f  max/warp %.1f  runs/lane %.2f" % (v[8] / max(v[0], 1), v[9] / max(v[11], 1), v[7] / max(v[0], 1)))
