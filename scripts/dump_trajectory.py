"""dev tool: per-iteration states + correspondences of the BASELINE workload for one build of the library, to
compare two builds bit for bit:
    VISMA_B200_LIB=libA.so python scripts/dump_trajectory.py /tmp/a.npz; ... libB ... /tmp/b.npz
    python scripts/dump_trajectory.py --cmp /tmp/a.npz /tmp/b.npz"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if sys.argv[1] == "--cmp":
    a, b = np.load(sys.argv[2]), np.load(sys.argv[3])
    bad = 0
    for k in a.files:
        # correspondences must be identical; poses / rmse may differ in the last bits (summation grouping)
        same = np.array_equal(a[k], b[k]) if "corr" in k else np.allclose(a[k], b[k], rtol=1e-9, atol=1e-12)
        if not same:
            bad += 1
            d = a[k] != b[k]
            print("DIFF", k, int(d.sum()), "of", d.size, "first", np.argwhere(d)[:5].tolist())
    print("identical" if not bad else "%d arrays differ" % bad)
    sys.exit(1 if bad else 0)
from visma_b200 import registration as reg, synth
n_iter = int(os.environ.get("N_ITER", "10"))
d = synth.make_room_scene(2_000_000, 32, 50_000, source_seed=0)
scene = reg.Scene(reg.PointCloud(d["scene_xyz"], d["scene_nrm"]), 0.075, device=0)
batch = reg.Batch(scene, [reg.PointCloud(p, n) for p, n in d["sources"]])
out = {}
for name, est in (("plane", reg.TransformationEstimationPointToPlane()), ("p2p", reg.TransformationEstimationPointToPoint())):
    batch.set_problems(d["T_init"])
    for it in range(n_iter):
        batch.iterate(est, 0.075, 1)
        res = batch.results(want_corr=True)
        out["%s_T_%d" % (name, it)] = np.stack([r.transformation_ for r in res])
        out["%s_rmse_%d" % (name, it)] = np.array([r.inlier_rmse_ for r in res])
        out["%s_corr_%d" % (name, it)] = np.concatenate([r.correspondence_set_[:, 1] + 0 * b for b, r in enumerate(res)])
        out["%s_ncorr_%d" % (name, it)] = np.array([len(r.correspondence_set_) for r in res])
np.savez(sys.argv[1], **out)
print("wrote", sys.argv[1])
