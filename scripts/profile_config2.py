"""dev tool: cProfile of the config-2 per-object flow (where the Python harness spends its time around the library calls)."""
import cProfile, pstats, os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from visma_b200 import annotation, dataset, io3d, registration as reg
with tempfile.TemporaryDirectory() as tmp:
    ds = dataset.write_clutter_dataset(tmp, n_scene=2_000_000, n_objects=8)
    cfg = io3d.load_json(ds["cfg_path"]); icp = cfg["ICP"]
    floor, _ = io3d.read_ply(os.path.join(ds["fragment_dir"], "floor.ply"))
    T0 = annotation.GravityAlignment(floor)
    objs = []
    for name in ds["entries"]:
        scan, _ = io3d.read_ply(os.path.join(ds["fragment_dir"], name + ".ply"))
        V, F = io3d.read_obj(os.path.join(ds["cad_dir"], name[:name.rfind("_")] + ".obj"))
        objs.append((name, scan, V, F))
    def run():
        ts = []
        for k, (name, scan, V, F) in enumerate(objs):
            t0 = time.perf_counter()
            n_scan = len(reg.VoxelDownSample(scan, icp["voxel_size"], 0).points_)
            model = reg.SamplePointCloudFromMesh(V, F, 2 * n_scan, seed=k, device=0)
            Ttot, info = annotation.AnnotateObject(scan, model, T0, icp, 0)
            ts.append(time.perf_counter() - t0)
        return ts
    run()
    pr = cProfile.Profile(); pr.enable(); ts = run(); pr.disable()
    print("ms per object %.2f" % (np.mean(ts) * 1e3))
    pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
