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
buf = (C.c_ulonglong * 16)()
L.vb200_debug_stats(None, 1)
names = ["valid", "prior", "", "coop", "coop:reach", "coop:overflow", "coop:nohome", "runs", "lane_steps", "warp_max_steps",
         "warps_w_coop", "warps", "hard", "hard:noprior", "hard:nosec"]
est = reg.TransformationEstimationPointToPlane()
for it in range(n_iter):
    batch.iterate(est, 0.075, 1)
    scene.sync()
    L.vb200_debug_stats(buf, 1)
    v = list(buf)
    print(it, " ".join("%s=%d" % (n, v[i]) for i, n in enumerate(names) if n),
          "| steps/lane %.1f  max/warp %.1f  runs/lane %.2f" % (v[8] / max(v[0], 1), v[9] / max(v[11], 1), v[7] / max(v[0], 1)))
