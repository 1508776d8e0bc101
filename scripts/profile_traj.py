"""dev tool for ncu: the BASELINE point-to-plane trajectory (2 M-point scene, 50k-point chair fragments), N_ITER
iterations from the initial poses, for the objects b with b % STRIDE == 0 (STRIDE=8: one rank's share at 8 GPUs).
    N_ITER=30 STRIDE=1 ncu --metrics gpu__time_duration.sum ... python scripts/profile_traj.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from visma_b200 import registration as reg, synth
n_iter = int(os.environ.get("N_ITER", "30"))
stride = int(os.environ.get("STRIDE", "1"))
d = synth.make_room_scene(2_000_000, 32, 50_000, source_seed=0)
scene = reg.Scene(reg.PointCloud(d["scene_xyz"], d["scene_nrm"]), 0.075, device=0)
mine = list(range(0, 32, stride))
batch = reg.Batch(scene, [reg.PointCloud(*d["sources"][b]) for b in mine])
batch.set_problems(np.stack([d["T_init"][b] for b in mine]))
batch.iterate(reg.TransformationEstimationPointToPlane(), 0.075, n_iter)
res = batch.results()
print("objects", len(mine), "iterations", n_iter, "fitness min %.4f" % min(r.fitness_ for r in res))
