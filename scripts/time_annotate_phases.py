"""dev tool: where AnnotateObject (feh::AnnotationTool's per-object flow, 24 yaw starts) spends its time on one fragment of the
generated clutter1 stand-in."""
import os, sys, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from visma_b200 import annotation, dataset, io3d, registration as reg
with tempfile.TemporaryDirectory() as tmp:
    ds = dataset.write_clutter_dataset(tmp, n_scene=2_000_000, n_objects=2)
    cfg = io3d.load_json(ds["cfg_path"]); icp = cfg["ICP"]
    floor, _ = io3d.read_ply(os.path.join(ds["fragment_dir"], "floor.ply"))
    T0 = annotation.GravityAlignment(floor)
    name = ds["entries"][0]
    scan, _ = io3d.read_ply(os.path.join(ds["fragment_dir"], name + ".ply"))
    V, F = io3d.read_obj(os.path.join(ds["cad_dir"], name[:name.rfind("_")] + ".obj"))
    for rep in range(4):
        t = [time.perf_counter()]
        sd = reg.VoxelDownSample(scan, icp["voxel_size"], 0).points_; t.append(time.perf_counter())
        model = reg.SamplePointCloudFromMesh(V, F, 2 * len(sd), seed=0, device=0); t.append(time.perf_counter())
        s2 = annotation._apply(T0, sd); t.append(time.perf_counter())
        scene = reg.Scene(s2, icp["distance_threshold"], 0); t.append(time.perf_counter())
        out = reg.RegisterModelToScene(model, scene, 24, icp["distance_threshold"], False); t.append(time.perf_counter())
        scene.close(); t.append(time.perf_counter())
        print("scan %d -> %d pts, model %d | voxel %.2f  sample %.2f  host %.2f  scene_create %.2f  register(24) %.2f  close %.2f ms"
              % ((len(scan), len(sd), len(model)) + tuple(1e3 * (b - a) for a, b in zip(t, t[1:]))))
