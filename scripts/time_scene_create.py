"""dev tool: vb200_scene_create from PAGEABLE host arrays (what a C++ caller's std::vector is) at a few sizes, and the
result checked against a scene built from pinned arrays (KNN answers equal)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from visma_b200 import registration as reg, synth
for n in (1_000_000, 2_000_000, 5_000_000):
    d = synth.make_room_scene(n, 1, 1000, seed=3)
    xyz, nrm = np.ascontiguousarray(d["scene_xyz"]), np.ascontiguousarray(d["scene_nrm"])
    ts = []
    for rep in range(6):
        t0 = time.perf_counter()
        sc = reg.Scene(reg.PointCloud(xyz, nrm), 0.075, device=0)
        sc.sync()
        ts.append(time.perf_counter() - t0)
        if rep < 5: sc.close()
    pin = lambda a: torch.from_numpy(a).pin_memory().numpy()
    t0 = time.perf_counter(); sp = reg.Scene(reg.PointCloud(pin(xyz), pin(nrm)), 0.075, device=0); sp.sync(); tp = time.perf_counter() - t0
    q = xyz[:: max(n // 5000, 1)] + 0.003
    ia, da = sc.SearchHybrid1(q, 0.075); ib, db = sp.SearchHybrid1(q, 0.075)
    print("n=%d  pageable %.2f ms (min of 5 after the first: %.2f)  pinned %.2f ms  %.1f GB/s  same answers: %s"
          % (n, ts[0] * 1e3, min(ts[1:]) * 1e3, tp * 1e3, 48e-9 * n / min(ts[1:]), bool((ia == ib).all() and (da == db).all())))
    sc.close(); sp.close()
