"""Small workloads touching every kernel, for compute-sanitizer (memcheck / racecheck / initcheck):
   compute-sanitizer --tool memcheck python scripts/sanitize_smoke.py"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from visma_b200 import registration as reg, renderer, synth

d = synth.make_room_scene(30000, 2, 1500, seed=3)
sc = reg.Scene(reg.PointCloud(d["scene_xyz"], d["scene_nrm"]), 0.075)
cl = [reg.PointCloud(p, n) for p, n in d["sources"]]
for est in (reg.TransformationEstimationPointToPlane(), reg.TransformationEstimationPointToPoint(),
            reg.TransformationEstimationPointToPlaneGravity()):
    r = reg.RegistrationICPBatch(cl, sc, 0.075, d["T_init"], est)
    print(type(est).__name__, [round(x.fitness_, 4) for x in r])
# a long run past convergence: the cached-neighbour test, two-candidate entries and near-empty part-B lists
b = reg.Batch(sc, cl)
b.set_problems(d["T_init"])
b.iterate(reg.TransformationEstimationPointToPlane(), 0.075, 25)
print("iterate", [round(x.fitness_, 4) for x in b.results()])
b.close()
i, d2 = sc.SearchHybrid1(synth.knn_queries(d["scene_xyz"], 3000), 0.075)
print("knn matched", int((i >= 0).sum()))
bi, bd2 = reg.SearchHybrid1BruteForce(d["scene_xyz"], synth.knn_queries(d["scene_xyz"], 3000), 0.075)
print("exhaustive knn equals grid", bool((bi == i).all() and (bd2 == d2).all()))
for nq in (1, 2, 3):
    reg.SearchHybrid1BruteForce(d["scene_xyz"], d["scene_xyz"][:nq] + 0.001, 0.075)
print("register", reg.RegisterModelToScene(cl[0], sc, 4, 0.05, True)["ncorr"])
print("estimate", reg.ComputeTransformation(reg.TransformationEstimationPointToPlane(), cl[0].points_,
                                            reg.PointCloud(d["scene_xyz"], d["scene_nrm"]),
                                            np.stack([np.arange(100), np.arange(100)], 1))[0, 0])
print("voxel", len(reg.VoxelDownSample(reg.PointCloud(d["scene_xyz"], d["scene_nrm"]), 0.05).points_))
V, F = synth.load_chair()
print("sample", reg.SamplePointCloudFromMesh(V, F, 2000, seed=1).shape)
ren = renderer.Renderer(120, 160)
ren.SetCamera(0.05, 10.0, 100.0, 100.0, 80.0, 60.0)
ren.SetMesh(V, F)
m = synth.make_T(np.eye(3), [0, 0, 0.6])
print("render", int((ren.RenderDepth(m) < 1).sum()), int(ren.RenderEdge(m).max()), int(ren.RenderMask(m).max()))
Vc, Fc = synth.cube_mesh()
ren.SetMesh(Vc, Fc)  # straddles the near plane and fills the image: clipped polygons + the queued big triangles
print("render cube", int((ren.RenderDepth(synth.make_T(synth.rot_xyz(0.3, 0.5, -0.2), [0, 0, 0.3])) < 1).sum()))
