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


def merge_statistics(a, b):
    """SampleStatistics::merge (common.h:268-310), vectorised over points: folds the statistics b of a second,
    independent batch of walks at the same points (e.g. the walks of each point split across ranks) into a."""
    def merge_moments(mean_a, m2_a, n_a, mean_b, m2_b, n_b):
        n = n_a + n_b
        f1, f2, fn = n_a.astype(np.float32), n_b.astype(np.float32), np.maximum(1, n).astype(np.float32)
        shape = (-1,) + (1,)*(mean_a.ndim - 1)
        f1, f2, fn = f1.reshape(shape), f2.reshape(shape), fn.reshape(shape)
        delta = mean_b - mean_a
        mean = np.where(f1 == 0, mean_b, mean_a + delta*(f2/fn))
        m2 = np.where(f1 == 0, m2_b, m2_a + m2_b + delta*delta*(f1*f2/fn))
        keep = (f2 == 0)
        return (np.where(keep, mean_a, mean).astype(np.float32), np.where(keep, m2_a, m2).astype(np.float32))

    out = abi.HostStatistics(a.n, a.dim)
    out.solutionMean, out.solutionM2 = merge_moments(a.solutionMean, a.solutionM2, a.nSolutionEstimates,
                                                     b.solutionMean, b.solutionM2, b.nSolutionEstimates)
    out.gradientMean, out.gradientM2 = merge_moments(a.gradientMean, a.gradientM2, a.nGradientEstimates,
                                                     b.gradientMean, b.gradientM2, b.nGradientEstimates)
    out.gradientMean = np.ascontiguousarray(out.gradientMean)
    out.gradientM2 = np.ascontiguousarray(out.gradientM2)
    out.nSolutionEstimates = a.nSolutionEstimates + b.nSolutionEstimates
    out.nGradientEstimates = a.nGradientEstimates + b.nGradientEstimates
    out.totalFirstSourceContribution = a.totalFirstSourceContribution + b.totalFirstSourceContribution
    out.totalDerivativeContribution = a.totalDerivativeContribution + b.totalDerivativeContribution
    out.totalWalkLength = a.totalWalkLength + b.totalWalkLength
    out.totalSplits = a.totalSplits + b.totalSplits
    return out


class DeviceStatistics:
    """The statistics of n points as ONE contiguous device buffer (float32 words; SoA sections in the order of the C ABI's
    zb_statistics: 2 + 2*dim + 2 float arrays, then 4 int32 arrays = 56 bytes per point in 3D), so that the gather after a
    sharded solve is a single all_gather_into_tensor on memory zb_solve_dev wrote — no host staging, no repacking."""

    def __init__(self, n, dim, device):
        import torch
        self.n, self.dim = n, dim
        self.words = 2 + 2*dim + 2 + 4
        self.buf = torch.zeros(self.words*n, dtype=torch.float32, device=device)
        self.sections = {}
        off = 0
        for k in STAT_FLOAT_KEYS:
            width = dim if k.startswith("gradient") else 1
            self.sections[k] = (off, width)
            off += width*n
        for k in STAT_INT_KEYS:
            self.sections[k] = (off, 1)
            off += n
        assert off == self.words*n

    def view(self, key, buf=None):
        """tensor view of one section of this buffer (or of one rank's slice of a gathered buffer)"""
        import torch
        b = self.buf if buf is None else buf
        off, width = self.sections[key]
        t = b[off:off + width*self.n]
        if key in STAT_INT_KEYS:
            t = t.view(torch.int32)
        return t.view(self.n, width) if width > 1 else t

    def as_struct(self):
        """abi.Statistics of device pointers into the buffer (what zb_solve_dev takes)"""
        base = self.buf.data_ptr()
        keys = ("solutionMean", "solutionM2", "gradientMean", "gradientM2", "totalFirstSourceContribution",
                "totalDerivativeContribution", "nSolutionEstimates", "nGradientEstimates", "totalWalkLength", "totalSplits")
        return abi.Statistics(*[base + 4*self.sections[k][0] for k in keys])

    def all_gather(self, out=None):
        """all ranks' buffers, [world, words*n] (rank r's points are shard_indices(n_total, r, world)); one collective"""
        import torch
        import torch.distributed as dist
        world = dist.get_world_size()
        if out is None:
            out = torch.empty((world, self.words*self.n), dtype=torch.float32, device=self.buf.device)
        dist.all_gather_into_tensor(out.view(-1), self.buf)
        return out

    def to_host(self, buf=None):
        """HostStatistics of this buffer (or of one rank's slice of a gathered buffer)"""
        st = abi.HostStatistics(self.n, self.dim)
        for k in STAT_FLOAT_KEYS + STAT_INT_KEYS:
            setattr(st, k, np.ascontiguousarray(self.view(k, buf).cpu().numpy()))
        return st


# ---- Reverse Walk on Stars and Boundary Value Caching over several GPUs ------------------------------------------
def allreduce_rws_totals(eval_pts, device="cpu"):
    """Reverse Walk on Stars with the SAMPLE walks sharded over the ranks and the evaluation points replicated: every rank's
    zb_rws_solve adds its samples' splats to its own copy of the six running sums per evaluation point
    (rws::EvaluationPoint, reverse_walk_solver.h:20-52); the sums of the whole job are their sum over the ranks — one
    all-reduce of 24 B per evaluation point (NCCL on the GPU box, gloo in the CPU test). eval_pts: abi.HostRwsEvalPoints,
    updated in place. Divide by the GLOBAL sample counts afterwards (allreduce_counts)."""
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(np.ascontiguousarray(eval_pts.totals)).to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    eval_pts.totals[...] = t.cpu().numpy()
    return eval_pts


def allreduce_counts(counts, device="cpu"):
    """the per-group sample counts (abi.HostRwsSamples.group_counts) of the whole job"""
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(np.asarray(counts, np.int64).copy()).to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


BVC_SAMPLE_KEYS = ("pt", "normal", "type", "pdf", "aligned", "dist_a", "dist_r", "robin_coeff", "solution", "normal_derivative", "contribution")


def allgather_bvc_cache(local, device="cpu"):
    """Boundary Value Caching with the CACHE samples sharded over the ranks for the walks (compute_sample_estimates on each
    rank's shard) and the evaluation points sharded for the splat: the splat needs every cache sample on every rank, so the
    finished shards are all-gathered — a few MB (boundary_value_caching.h: the cache is 10^4..10^5 samples). local: a
    bvc.Samples; returns the bvc.Samples of all ranks in rank order. The pdf of a sample is that of the whole sampler, so
    nothing is rescaled."""
    import torch
    import torch.distributed as dist
    from . import bvc
    world = dist.get_world_size()
    n = torch.tensor([len(local)], dtype=torch.int64, device=device)
    sizes = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(sizes, n)
    sizes = [int(s.item()) for s in sizes]
    pad = max(sizes) if sizes else 0
    parts = [dict() for _ in range(world)]
    for k in BVC_SAMPLE_KEYS:
        a = np.ascontiguousarray(getattr(local, k))
        width = int(np.prod(a.shape[1:])) if a.ndim > 1 else 1
        buf = np.zeros((pad, width), a.dtype)
        buf[:len(a)] = a.reshape(len(a), width)
        t = torch.from_numpy(buf).to(device)
        out = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        for r in range(world):
            g = out[r].cpu().numpy()[:sizes[r]]
            parts[r][k] = g.reshape((sizes[r],) + a.shape[1:])
    gathered = []
    for r in range(world):
        if sizes[r] == 0:
            continue
        s = bvc.Samples(parts[r]["pt"], parts[r]["normal"], 0, 1.0, False)
        for k in BVC_SAMPLE_KEYS[2:]:
            setattr(s, k, parts[r][k])
        gathered.append(s)
    return bvc.Samples.concatenate(gathered)
