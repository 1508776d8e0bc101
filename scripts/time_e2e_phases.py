"""dev tool: where one vb200_icp_run-equivalent call spends its host time (BASELINE workload, pinned sources)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from visma_b200 import registration as reg, synth
d = synth.make_room_scene(2_000_000, 32, 50_000, source_seed=0)
scene = reg.Scene(reg.PointCloud(d["scene_xyz"], d["scene_nrm"]), 0.075, device=0)
src = torch.from_numpy(np.concatenate([p for p, n in d["sources"]])).pin_memory().numpy()
clouds = [reg.PointCloud(src[i * 50000:(i + 1) * 50000], d["sources"][i][1]) for i in range(32)]
est = reg.TransformationEstimationPointToPlane()
for crit in (reg.ICPConvergenceCriteria(0.0, 0.0, 30), reg.ICPConvergenceCriteria()):
    acc = np.zeros(5)
    for rep in range(12):
        t = [time.perf_counter()]
        b = reg.Batch(scene, clouds); t.append(time.perf_counter())
        b.set_problems(d["T_init"]); t.append(time.perf_counter())
        b.run(est, 0.075, crit); t.append(time.perf_counter())
        r = b.results(); t.append(time.perf_counter())
        b.close(); t.append(time.perf_counter())
        if rep >= 2:
            acc += np.diff(t)
    print("max_iter=%d rel=%g:" % (crit.max_iteration_, crit.relative_fitness_),
          "create(H2D+sort) %.2f  set_problems %.2f  run(enqueue) %.2f  results(wait+D2H) %.2f  destroy %.2f  ms"
          % tuple(acc / 10 * 1e3), "iters", max(x.iterations_ for x in r))
