"""The N > 1 path on real GPUs: world_size-2 processes each run THEIR share of the objects through the CUDA
library (shard.register_sharded -> vb200_icp_run) and all-gather the pose table, which must equal the single-rank
RegistrationICPBatch of all the objects bit for bit (each object's solve is independent of what else is in the
batch: SURVEY §8e, src/annotation.cpp:103-141).  NCCL when the box has two devices, otherwise both ranks share
cuda:0 and the 4 KB table goes through gloo (NCCL refuses two ranks on one device)."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, small_scene

pytestmark = pytest.mark.gpu

N_OBJECTS, SCENE_KW = 5, dict(n_scene=120000, n_objects=5, m=5000, seed=31)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist
    from visma_b200 import registration as reg, shard
    ndev = torch.cuda.device_count()
    use_nccl = ndev >= world
    device = rank if use_nccl else 0
    torch.cuda.set_device(device)
    if use_nccl:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", device))
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    d = small_scene(**SCENE_KW)
    scene = reg.Scene(reg.PointCloud(d["scene_xyz"], d["scene_nrm"]), 0.075, device=device)
    sources = [reg.PointCloud(p, n) for p, n in d["sources"]]
    table = shard.register_sharded(scene, sources, d["T_init"], 0.075, reg.TransformationEstimationPointToPlane(),
                                   device=torch.device("cuda", device) if use_nccl else None)
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, table, "nccl" if use_nccl else "gloo"))


def test_register_sharded_equals_single_rank_bit_for_bit(vb):
    import torch.multiprocessing as mp
    from visma_b200 import shard
    d = small_scene(**SCENE_KW)
    scene = vb.reg.Scene(vb.reg.PointCloud(d["scene_xyz"], d["scene_nrm"]), 0.075)
    sources = [vb.reg.PointCloud(p, n) for p, n in d["sources"]]
    res = vb.reg.RegistrationICPBatch(sources, scene, 0.075, d["T_init"], vb.reg.TransformationEstimationPointToPlane(),
                                      want_corr=False)
    single = shard.unpack_table(shard.pack_results(res, N_OBJECTS, 0, 1)[None], N_OBJECTS, 1)
    assert all(r.fitness_ > 0.9 for r in res)
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + 911
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, table, backend in got:
        assert table.shape == (N_OBJECTS, shard.ROW)
        assert np.array_equal(table, single), (rank, backend, np.abs(table - single).max())


def _global_case():
    """ICPRefinement's scene_est (src/evaluation.cpp:244-274): the union of the posed model samples, one global
    transform to estimate from a small perturbation."""
    from visma_b200 import synth
    d = small_scene(**SCENE_KW)
    pts = np.concatenate([p @ T[:3, :3].T + T[:3, 3] for (p, _), T in zip(d["sources"], d["T_gt"])])
    nrm = np.concatenate([n @ T[:3, :3].T for (_, n), T in zip(d["sources"], d["T_gt"])])
    init = synth.make_T(synth.rot_xyz(0.004, -0.01, 0.003), [0.01, -0.004, 0.012])
    return d, pts, nrm, init


def _global_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist
    from visma_b200 import registration as reg, shard
    use_nccl = torch.cuda.device_count() >= world
    device = rank if use_nccl else 0
    torch.cuda.set_device(device)
    if use_nccl:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", device))
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    d, pts, nrm, init = _global_case()
    scene = reg.Scene(reg.PointCloud(d["scene_xyz"], d["scene_nrm"]), 0.075, device=device)
    mine = slice(rank * len(pts) // world, (rank + 1) * len(pts) // world)
    res = shard.register_global_sharded(scene, reg.PointCloud(pts[mine], nrm[mine]), init, 0.075,
                                        reg.TransformationEstimationPointToPlane(),
                                        device=torch.device("cuda", device))
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, np.asarray(res.transformation_), res.fitness_, res.inlier_rmse_, res.iterations_))


def test_one_cloud_sharded_over_ranks_equals_single_gpu(vb):
    """"Next" row 3: ONE source cloud split over two ranks, the per-iteration totals all-reduced
    (vb200_batch_pass -> all-reduce of 256 bytes -> vb200_batch_solve).  Every rank must hold the identical
    transform, equal to the single-GPU alignment of the whole cloud up to the summation order of the totals."""
    import torch.multiprocessing as mp
    from visma_b200 import synth
    d, pts, nrm, init = _global_case()
    scene = vb.reg.Scene(vb.reg.PointCloud(d["scene_xyz"], d["scene_nrm"]), 0.075)
    whole = vb.reg.RegistrationICP(vb.reg.PointCloud(pts, nrm), scene, 0.075, init,
                                   vb.reg.TransformationEstimationPointToPlane())
    assert whole.fitness_ > 0.9
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + 977
    procs = [ctx.Process(target=_global_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert np.array_equal(got[0][1], got[1][1]) and got[0][2:] == got[1][2:], "ranks diverged"
    rot, tr = synth.pose_error(got[0][1], whole.transformation_)
    assert rot < 1e-9 and tr < 1e-9, (rot, tr)
    assert abs(got[0][2] - whole.fitness_) < 1e-5 and abs(got[0][3] - whole.inlier_rmse_) < 1e-9
    assert abs(got[0][4] - whole.iterations_) <= 1
