"""Phase timing of one vb200_icp_run-equivalent on the BASELINE workload (dev tool, run on the GPU box)."""
import sys, time, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from visma_b200 import registration as reg, synth

d = synth.make_room_scene(2_000_000, 32, 50_000)
t = time.perf_counter(); scene = reg.Scene(reg.PointCloud(d["scene_xyz"], d["scene_nrm"]), 0.075); print("scene_create %.1f ms" % ((time.perf_counter() - t) * 1e3))
clouds = [reg.PointCloud(p, n) for p, n in d["sources"]]
est = reg.TransformationEstimationPointToPlane()
for rep in range(3):
    t0 = time.perf_counter(); b = reg.Batch(scene, clouds); t1 = time.perf_counter()
    b.set_problems(d["T_init"]); t2 = time.perf_counter()
    b.run(est, 0.075, reg.ICPConvergenceCriteria(0.0, 0.0, 30)); t3 = time.perf_counter()
    r = b.results(); t4 = time.perf_counter()
    b.close(); t5 = time.perf_counter()
    print("batch_create %.1f  set_problems %.1f  run(launch) %.1f  results(sync) %.1f  destroy %.1f ms" %
          tuple((y - x) * 1e3 for x, y in ((t0, t1), (t1, t2), (t2, t3), (t3, t4), (t4, t5))))
    t0 = time.perf_counter()
    reg.RegistrationICPBatch(clouds, scene, 0.075, d["T_init"], est, reg.ICPConvergenceCriteria(0.0, 0.0, 30), want_corr=False)
    print("RegistrationICPBatch total %.1f ms" % ((time.perf_counter() - t0) * 1e3))
