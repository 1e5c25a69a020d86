"""Multi-GPU sharding of the estimator: evaluation points are independent (walk_on_stars.h:159 "solve the PDE
at each point independently"), so each rank (one process per GPU) solves an interleaved shard of the points
against its own replica of the boundary BVHs. There is no data-path collective; the only communication is
the gather of the per-point SampleStatistics (<= 56 B/point) after the solve, over torch.distributed
(NCCL on the GPU box, gloo in the CPU tests)."""
import numpy as np

from . import abi

STAT_FLOAT_KEYS = ("solutionMean", "solutionM2", "gradientMean", "gradientM2", "totalFirstSourceContribution",
                   "totalDerivativeContribution")
STAT_INT_KEYS = ("nSolutionEstimates", "nGradientEstimates", "totalWalkLength", "totalSplits")


def shard_indices(n, rank, world):
    """interleaved shard: neighbouring points (similar walk lengths) land on different ranks -> load balance"""
    return np.arange(rank, n, world, dtype=np.int64)


def shard_size(n, world):
    return (n + world - 1)//world


def pack_statistics(stats, pad_to):
    """HostStatistics -> one float32 [pad_to, F] and one int32 [pad_to, 4] array"""
    f = np.concatenate([np.asarray(getattr(stats, k), np.float32).reshape(stats.n, -1) for k in STAT_FLOAT_KEYS], axis=1)
    i = np.stack([np.asarray(getattr(stats, k), np.int32) for k in STAT_INT_KEYS], axis=1)
    fp = np.zeros((pad_to, f.shape[1]), np.float32)
    ip = np.zeros((pad_to, 4), np.int32)
    fp[:stats.n] = f
    ip[:stats.n] = i
    return fp, ip


def unpack_statistics(fp, ip, n, dim):
    st = abi.HostStatistics(n, dim)
    col = 0
    for k in STAT_FLOAT_KEYS:
        width = dim if k.startswith("gradient") else 1
        block = fp[:n, col:col + width]
        setattr(st, k, np.ascontiguousarray(block if width > 1 else block[:, 0]))
        col += width
    for j, k in enumerate(STAT_INT_KEYS):
        setattr(st, k, np.ascontiguousarray(ip[:n, j]))
    return st


def gather_statistics(local_stats, n_total, dim, rank, world, device="cpu"):
    """all-gathers the shards' statistics and returns the HostStatistics of all n_total points in their
    original order (on every rank). Requires an initialised torch.distributed process group when world > 1."""
    if world == 1:
        return local_stats
    import torch
    import torch.distributed as dist
    pad = shard_size(n_total, world)
    fp, ip = pack_statistics(local_stats, pad)
    tf = torch.from_numpy(fp).to(device)
    ti = torch.from_numpy(ip).to(device)
    of = torch.empty((world*pad, fp.shape[1]), dtype=torch.float32, device=device)
    oi = torch.empty((world*pad, 4), dtype=torch.int32, device=device)
    dist.all_gather_into_tensor(of, tf)
    dist.all_gather_into_tensor(oi, ti)
    of = of.cpu().numpy().reshape(world, pad, -1)
    oi = oi.cpu().numpy().reshape(world, pad, 4)
    full_f = np.zeros((n_total, fp.shape[1]), np.float32)
    full_i = np.zeros((n_total, 4), np.int32)
    for r in range(world):
        idx = shard_indices(n_total, r, world)
        full_f[idx] = of[r, :len(idx)]
        full_i[idx] = oi[r, :len(idx)]
    return unpack_statistics(full_f, full_i, n_total, dim)
