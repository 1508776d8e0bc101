"""dev tool: the 30-iteration BASELINE trajectory of ONE rank's share at world size W, on one GPU (what bench.py times per
step under torchrun, without the other ranks): CUDA events on the library's stream, 256 MiB L2 flush before every step.
    W=8 VISMA_B200_LIB=build/variants/lib_x.so python scripts/time_shard_traj.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from visma_b200 import registration as reg, shard, synth
W = int(os.environ.get("W", "8"))
reps = int(os.environ.get("REPS", "3"))
d = synth.make_room_scene(2_000_000, 32, 50_000, source_seed=0)
dev = torch.device("cuda", 0)
scene = reg.Scene(reg.PointCloud(d["scene_xyz"], d["scene_nrm"]), 0.075, device=0)
mine = shard.shard_objects(32, 0, W)
batch = reg.Batch(scene, [reg.PointCloud(*d["sources"][b]) for b in mine])
T_init = np.ascontiguousarray(d["T_init"][mine])
est = reg.TransformationEstimationPointToPlane()
stream = torch.cuda.ExternalStream(scene.stream(), device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def traj():
    batch.set_problems(T_init)
    out = []
    for it in range(30):
        flush.fill_(1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        batch.iterate(est, 0.075, 1)
        e1.record(stream)
        torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1))
    return np.array(out)
traj()
t = np.min([traj() for _ in range(reps)], axis=0)
print("W=%d objects=%d  total %.4f ms  first6 %.4f  last10 mean %.4f  it/s %.0f" % (W, len(mine), t.sum(), t[:6].sum(), t[20:].mean(), 30 / (t.sum() * 1e-3)))
print("   ", [round(float(x), 4) for x in t])
