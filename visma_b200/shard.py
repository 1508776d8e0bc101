"""Multi-GPU sharding of the ICP path: one process per GPU, objects round-robin over ranks, the scene
replicated, and ONE collective — an all-gather of the final pose table.  Each object's solve is
independent (src/annotation.cpp:103-141), so there is no data-path exchange to fuse with a kernel.

torch.distributed is plumbing only (NCCL over NVLink on the GPUs, gloo in the CPU tests).
"""
import numpy as np

ROW = 20  # T (16, row-major) + fitness + rmse + ncorr + iterations


def shard_objects(n_objects, rank, world):
    """Objects owned by `rank`: {b : b mod world == rank}.  All yaw initialisations of an object stay on its
    rank so RegisterModelToScene's arg-max (src/annotation.cpp:59-61) is local."""
    return list(range(rank, n_objects, world))


def rows_per_rank(n_objects, world):
    return (n_objects + world - 1) // world


def pack_results(results, n_objects, rank, world):
    """results: list of RegistrationResult for shard_objects(...) in order -> padded [rows_per_rank, ROW]."""
    out = np.zeros((rows_per_rank(n_objects, world), ROW), np.float64)
    for k, r in enumerate(results):
        out[k, :16] = np.asarray(r.transformation_).reshape(-1)
        out[k, 16] = r.fitness_
        out[k, 17] = r.inlier_rmse_
        out[k, 18] = len(r.correspondence_set_)
        out[k, 19] = r.iterations_
    return out


def unpack_table(gathered, n_objects, world):
    """gathered: [world, rows_per_rank, ROW] -> [n_objects, ROW] in object order."""
    table = np.zeros((n_objects, ROW), np.float64)
    for r in range(world):
        for k, b in enumerate(shard_objects(n_objects, r, world)):
            table[b] = gathered[r, k]
    return table


def all_gather_poses(local_rows, n_objects, device=None, group=None):
    """The path's single collective.  local_rows: this rank's padded [rows_per_rank, ROW] table.
    Returns the full [n_objects, ROW] table on every rank."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    t = torch.from_numpy(np.ascontiguousarray(local_rows))
    if device is not None:
        t = t.to(device)
    if world == 1:
        return unpack_table(t.cpu().numpy()[None], n_objects, 1)
    # one flat collective into one tensor (the list form of all_gather costs a copy per rank: 0.1 ms for 4 KB)
    out = torch.empty((world * t.shape[0], t.shape[1]), dtype=t.dtype, device=t.device)  # ranks concatenated along dim 0
    dist.all_gather_into_tensor(out, t, group=group)
    return unpack_table(out.cpu().numpy().reshape(world, t.shape[0], t.shape[1]), n_objects, world)


def register_sharded(scene, sources, inits, max_dist, estimation, criteria=None, device=None, group=None):
    """Run this rank's share of a global list of ICP problems and all-gather the pose table.
    sources / inits are the GLOBAL lists (every rank sees the same ones); returns [n_objects, ROW]."""
    import torch.distributed as dist
    from . import registration as reg
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    mine = shard_objects(len(sources), rank, world)
    res = reg.RegistrationICPBatch([sources[b] for b in mine], scene, max_dist,
                                   np.asarray([inits[b] for b in mine]).reshape(-1, 4, 4), estimation, criteria,
                                   want_corr=False) if mine else []
    return all_gather_poses(pack_results(res, len(sources), rank, world), len(sources), device, group)


def register_global_sharded(scene, source_shard, init, max_dist, estimation, criteria=None, device=None,
                            group=None):
    """ONE ICP problem whose source cloud is sharded over the ranks (ICPRefinement's global transform,
    src/evaluation.cpp:244-274): every rank searches its shard against its scene replica, the 32 per-problem
    totals are all-reduced (the path's one per-iteration collective: 256 bytes), and every rank applies the
    identical estimator update.  Returns the RegistrationResult (same on every rank)."""
    import torch
    import torch.distributed as dist
    from . import registration as reg
    criteria = criteria or reg.ICPConvergenceCriteria()
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    dev = device if device is not None else torch.device("cuda", scene.device)
    batch = reg.Batch(scene, [source_shard])
    batch.set_problems(np.asarray(init, np.float64).reshape(1, 16))
    totals = torch.zeros(32, dtype=torch.float64, device=dev)
    batch.set_totals_buffer(totals.data_ptr())
    # NCCL reduces the device buffer in place; under gloo (two ranks sharing one GPU in the tests) the 256 bytes go
    # through the host — either way every rank ends up with bit-identical totals
    via_host = world > 1 and dist.get_backend(group) != "nccl"

    def all_reduce(t):
        if via_host:
            h = t.cpu()
            dist.all_reduce(h, group=group)
            t.copy_(h)
        else:
            dist.all_reduce(t, group=group)

    n_local = torch.tensor([batch.sizes[0]], dtype=torch.int64, device=dev)
    if world > 1:
        all_reduce(n_local)
    n_global = [int(n_local.item())]
    for it in range(criteria.max_iteration_ + 1):
        batch.pass_(estimation, max_dist)
        scene.sync()                      # totals written on the library's stream
        if world > 1:
            all_reduce(totals)
            torch.cuda.synchronize(dev)
        batch.solve(estimation, max_dist, criteria, it, n_global)
    res = batch.results()[0]
    batch.close()
    return res
