"""dev tool for ncu: ONE object's RegisterModelToScene of the config-2 stand-in (24 yaw starts over one model cloud)."""
import os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from visma_b200 import annotation, dataset, io3d, registration as reg
with tempfile.TemporaryDirectory() as tmp:
    ds = dataset.write_clutter_dataset(tmp, n_scene=2_000_000, n_objects=8)
    cfg = io3d.load_json(ds["cfg_path"]); icp = cfg["ICP"]
    floor, _ = io3d.read_ply(os.path.join(ds["fragment_dir"], "floor.ply"))
    T0 = annotation.GravityAlignment(floor)
    name = ds["entries"][int(os.environ.get("OBJ", "0"))]
    scan, _ = io3d.read_ply(os.path.join(ds["fragment_dir"], name + ".ply"))
    V, F = io3d.read_obj(os.path.join(ds["cad_dir"], name[:name.rfind("_")] + ".obj"))
    for rep in range(int(os.environ.get("REPS", "1"))):
        t0 = time.perf_counter()
        n_scan = len(reg.VoxelDownSample(scan, icp["voxel_size"], 0).points_)
        model = reg.SamplePointCloudFromMesh(V, F, 2 * n_scan, seed=0, device=0)
        Ttot, info = annotation.AnnotateObject(scan, model, T0, icp, 0)
        print("object", name, "scan", n_scan, "model", len(getattr(model, "points_", model)), "%.2f ms" % ((time.perf_counter() - t0) * 1e3), info if rep == 0 else "")
