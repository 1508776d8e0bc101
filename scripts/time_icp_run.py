"""dev tool: vb200_icp_run on the BASELINE workload with pinned packed sources (what bench.py's e2e leg calls):
median ms per call for 30 forced iterations and for the default criteria."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from visma_b200 import registration as reg, synth
d = synth.make_room_scene(2_000_000, 32, 50_000, source_seed=0)
scene = reg.Scene(reg.PointCloud(d["scene_xyz"], d["scene_nrm"]), 0.075, device=0)
src = torch.from_numpy(np.concatenate([p for p, n in d["sources"]])).pin_memory()
packed = (src.numpy(), np.arange(33, dtype=np.int64) * 50_000, True)
est = reg.TransformationEstimationPointToPlane()
for name, crit in (("30 iterations", reg.ICPConvergenceCriteria(0.0, 0.0, 30)), ("default criteria", reg.ICPConvergenceCriteria())):
    ts = []
    for rep in range(14):
        t0 = time.perf_counter()
        r = reg.RegistrationICPBatch(None, scene, 0.075, d["T_init"], est, crit, want_corr=False, packed=packed)
        ts.append(time.perf_counter() - t0)
    ts = np.array(ts[3:]) * 1e3
    print("%-18s median %.2f ms  min %.2f  max %.2f  iters %d-%d" % (name, np.median(ts), ts.min(), ts.max(),
          min(x.iterations_ for x in r), max(x.iterations_ for x in r)))
