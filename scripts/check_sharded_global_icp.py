"""Multi-GPU check of shard.register_global_sharded (run under torchrun on >= 2 GPUs):
   torchrun --nproc-per-node 2 scripts/check_sharded_global_icp.py
Every rank aligns its slice of ONE 200k-point source cloud against the replicated scene with one 256-byte
all-reduce per iteration; the result must equal the single-GPU alignment of the whole cloud."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch
import torch.distributed as dist
from visma_b200 import registration as reg, shard, synth

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
d = synth.make_room_scene(1_000_000, 8, 25_000, seed=31)
scene = reg.Scene(reg.PointCloud(d["scene_xyz"], d["scene_nrm"]), 0.075, device=local)
# ICPRefinement's scene_est: the union of the posed model samples, one global transform to estimate
est_pts = np.concatenate([p @ T[:3, :3].T + T[:3, 3] for (p, _), T in zip(d["sources"], d["T_gt"])])
est_nrm = np.concatenate([n @ T[:3, :3].T for (_, n), T in zip(d["sources"], d["T_gt"])])
off = synth.make_T(synth.rot_xyz(0.004, -0.01, 0.003), [0.01, -0.004, 0.012])
estimation = reg.TransformationEstimationPointToPlane()
mine = slice(rank * len(est_pts) // world, (rank + 1) * len(est_pts) // world)
t0 = time.perf_counter()
res = shard.register_global_sharded(scene, reg.PointCloud(est_pts[mine], est_nrm[mine]), off, 0.075, estimation)
dt = time.perf_counter() - t0
if rank == 0:
    whole = reg.RegistrationICP(reg.PointCloud(est_pts, est_nrm), scene, 0.075, off, estimation)
    rot, tr = synth.pose_error(res.transformation_, whole.transformation_)
    print("world", world, "sharded vs single-GPU: rot %.2e rad, trans %.2e m, fitness %.6f vs %.6f, iters %d vs %d, %.1f ms"
          % (rot, tr, res.fitness_, whole.fitness_, res.iterations_, whole.iterations_, dt * 1e3))
    assert rot < 1e-9 and tr < 1e-9 and abs(res.fitness_ - whole.fitness_) < 1e-5
T = torch.tensor(res.transformation_, device="cuda")
Ts = [torch.empty_like(T) for _ in range(world)]
dist.all_gather(Ts, T)
assert all(torch.equal(Ts[0], t) for t in Ts), "ranks diverged"
if rank == 0:
    print("all ranks hold the identical transform: OK")
dist.destroy_process_group()
