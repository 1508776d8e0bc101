"""The kernels added or changed in round 1's third session, for compute-sanitizer (memcheck / racecheck / initcheck):
the warp-flattened rasteriser (+ clipped / queued triangles), the exhaustive TMA-staged KNN, the breadth-first
warp-per-query search, and one ICP run (k_solve's reduction).  scripts/sanitize_smoke.py covers everything else."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from visma_b200 import registration as reg, renderer, synth

d = synth.make_room_scene(30000, 2, 1500, seed=3)
sc = reg.Scene(reg.PointCloud(d["scene_xyz"], d["scene_nrm"]), 0.075)
cl = [reg.PointCloud(p, n) for p, n in d["sources"]]
r = reg.RegistrationICPBatch(cl, sc, 0.075, d["T_init"], reg.TransformationEstimationPointToPlane())
print("icp", [round(x.fitness_, 4) for x in r])
q = np.concatenate([synth.knn_queries(d["scene_xyz"], 3000), np.random.default_rng(0).uniform(-1, 7, (500, 3))])
i, d2 = sc.SearchHybrid1(q, 0.075)
bi, bd2 = reg.SearchHybrid1BruteForce(d["scene_xyz"], q, 0.075)
print("knn matched", int((i >= 0).sum()), "exhaustive equals grid", bool((bi == i).all() and (bd2 == d2).all()))
for nq in (1, 2, 3):
    reg.SearchHybrid1BruteForce(d["scene_xyz"][:2500], d["scene_xyz"][:nq] + 0.001, 0.075)
V, F = synth.load_chair()
ren = renderer.Renderer(120, 160)
ren.SetCamera(0.05, 10.0, 100.0, 100.0, 80.0, 60.0)
ren.SetMesh(V, F)
m = synth.make_T(np.eye(3), [0, 0, 0.6])
print("render", int((ren.RenderDepth(m) < 1).sum()), int(ren.RenderEdge(m).max()), int(ren.RenderMask(m).max()))
Vc, Fc = synth.cube_mesh()
ren.SetMesh(Vc, Fc)  # straddles the near plane and fills the image: clipped polygons + the queued big triangles
print("render cube", int((ren.RenderDepth(synth.make_T(synth.rot_xyz(0.3, 0.5, -0.2), [0, 0, 0.3])) < 1).sum()))
