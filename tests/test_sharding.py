"""CPU test (gloo, world_size 2) of the multi-GPU path's host logic: interleaved sharding of the evaluation
points + gather of the per-point statistics reproduces the single-process result exactly. The per-shard solver
here is the test oracle (the compiled reference) because there is no GPU in this container; on the GPU box the
same sharding code drives zb_solve on each rank (bench.py --gpus N)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests import cases
from zombie_b200 import abi, sharding


def test_shards_partition_the_points():
    for n, world in ((10, 1), (10, 3), (1 << 20, 8), (7, 8)):
        allidx = np.concatenate([sharding.shard_indices(n, r, world) for r in range(world)])
        assert sorted(allidx.tolist()) == list(range(n))
        assert max(len(sharding.shard_indices(n, r, world)) for r in range(world)) == sharding.shard_size(n, world)


def test_pack_unpack_roundtrip():
    st = abi.HostStatistics(5, 3)
    rng = np.random.RandomState(0)
    for k in sharding.STAT_FLOAT_KEYS:
        getattr(st, k)[...] = rng.normal(size=getattr(st, k).shape)
    for k in sharding.STAT_INT_KEYS:
        getattr(st, k)[...] = rng.randint(0, 100, size=5)
    fp, ip = sharding.pack_statistics(st, 8)
    back = sharding.unpack_statistics(fp, ip, 5, 3)
    for k in sharding.STAT_FLOAT_KEYS + sharding.STAT_INT_KEYS:
        assert np.array_equal(getattr(st, k), getattr(back, k))


def _solve(points, idx, n_walks):
    from oracle import ref
    (pos, tri), absorbing, reflecting = cases.mesh3d(2)
    rs = ref.RefScene(3, **cases.scene_kwargs(3, abi.REFLECTING_NEUMANN, absorbing, reflecting))
    x = points[idx]
    _, da, _, _ = rs.closest_point(abi.ABSORBING, x)
    _, dr, _, _ = rs.closest_point(abi.REFLECTING, x)
    pts = abi.HostSamplePoints(x, da, dr, nWalks=n_walks, estimationQuantity=abi.SOLUTION, streams=idx.astype(np.uint64))
    st = abi.HostStatistics(len(idx), 3)
    rs.solve(abi.WALK_ON_STARS, cases.quadratic_pde(), abi.WalkSettings.default(), pts, st, threads=1)
    return st


def _worker(rank, world, port, n, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    points = cases.query_points(n, 3, 0.7, seed=3)
    idx = sharding.shard_indices(n, rank, world)
    st = _solve(points, idx, 8)
    full = sharding.gather_statistics(st, n, 3, rank, world, device="cpu")
    # the device-resident form of the same gather (bench.py --gpus N: one all_gather_into_tensor on the buffer zb_solve_dev
    # wrote), here on CPU tensors: every rank pads its shard to the common shard size
    pad = sharding.shard_size(n, world)
    ds = sharding.DeviceStatistics(pad, 3, "cpu")
    for k in sharding.STAT_FLOAT_KEYS + sharding.STAT_INT_KEYS:
        ds.view(k)[:len(idx)] = torch.from_numpy(np.ascontiguousarray(getattr(st, k)))
    gathered = ds.all_gather()
    assert gathered.shape == (world, ds.words*pad) and ds.words*4 == 56
    for r in range(world):
        ridx = sharding.shard_indices(n, r, world)
        part = ds.to_host(gathered[r])
        for k in sharding.STAT_FLOAT_KEYS + sharding.STAT_INT_KEYS:
            assert np.array_equal(getattr(part, k)[:len(ridx)], getattr(full, k)[ridx]), k
    if rank == 0:
        np.savez(out, mean=full.solutionMean, m2=full.solutionM2, n=full.nSolutionEstimates, wl=full.totalWalkLength)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_solve_equals_single_process(ref, tmp_path):
    n = 101   # ragged: the shards have 51 and 50 points
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path/"gathered.npz")
    mp.spawn(_worker, args=(2, port, n, out), nprocs=2, join=True)
    g = np.load(out)
    points = cases.query_points(n, 3, 0.7, seed=3)
    single = _solve(points, np.arange(n, dtype=np.int64), 8)
    assert np.array_equal(g["mean"], single.solutionMean)
    assert np.array_equal(g["m2"], single.solutionM2)
    assert np.array_equal(g["n"], single.nSolutionEstimates)
    assert np.array_equal(g["wl"], single.totalWalkLength)


# ---- Reverse Walk on Stars: samples sharded, totals all-reduced; BVC: cache shards all-gathered --------------------------------
def _rws_problem():
    from zombie_b200 import bvc, procedural as P
    whole, absorbing, reflecting = cases.mesh3d(2)
    sa, sr = bvc.UniformBoundarySampler(*absorbing, seed=1), bvc.UniformBoundarySampler(*reflecting, seed=2)
    sd = bvc.UniformDomainSampler(lambda p: P.bumpy_sphere_inside(p), [-1.2]*3, [1.2]*3, P.bumpy_sphere_volume(), seed=3)
    sa.initialize(0.01, False)
    sr.initialize(0.0, False)
    s = bvc.Samples.concatenate([sa.generate_samples(600, abi.ON_ABSORBING_BOUNDARY, 0.01), sr.generate_samples(300, abi.ON_REFLECTING_BOUNDARY, 0.0),
                                 sd.generate_samples(300)])
    return absorbing, reflecting, s, cases.query_points(40, 3, 0.6, seed=5)


def _rws_solve(s, idx, x):
    from oracle import ref
    absorbing, reflecting = _rws_problem()[:2]
    rs = ref.RefScene(3, **cases.scene_kwargs(3, abi.REFLECTING_NEUMANN, absorbing, reflecting))
    da = rs.closest_point(abi.ABSORBING, s.pt[idx], want_normal=False)[1]
    dr = rs.closest_point(abi.REFLECTING, s.pt[idx], want_normal=False)[1]
    smp = abi.HostRwsSamples(s.pt[idx], s.normal[idx], s.type[idx], s.pdf[idx], da, dr, seed=0xF00D, streams=idx.astype(np.uint64))
    ev = abi.HostRwsEvalPoints(x)
    rs.rws_solve(cases.quadratic_pde(), abi.WalkSettings.default(), smp, ev, 0.01)
    return smp, ev


def _rws_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    absorbing, reflecting, s, x = _rws_problem()
    idx = sharding.shard_indices(len(s), rank, world)
    smp, ev = _rws_solve(s, idx, x)
    sharding.allreduce_rws_totals(ev)
    counts = sharding.allreduce_counts(smp.group_counts())
    # the cache all-gather of Boundary Value Caching on the same shards: every rank ends up with all samples, in rank order
    from zombie_b200 import bvc
    local = bvc.Samples(s.pt[idx], s.normal[idx], 0, 1.0, False)
    for k in sharding.BVC_SAMPLE_KEYS[2:]:
        setattr(local, k, getattr(s, k)[idx])
    local.solution = (idx + 0.5).astype(np.float32)            # stands for the estimates a rank computed for its shard
    full = sharding.allgather_bvc_cache(local)
    order = np.concatenate([sharding.shard_indices(len(s), r, world) for r in range(world)])
    assert np.array_equal(full.pt, s.pt[order]) and np.array_equal(full.type, s.type[order]) and np.array_equal(full.pdf, s.pdf[order])
    assert np.array_equal(full.solution, (order + 0.5).astype(np.float32))
    if rank == 0:
        np.savez(out, totals=ev.totals, counts=counts)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_reverse_walks_allreduce_to_the_single_process_totals(ref, tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path/"rws.npz")
    mp.spawn(_rws_worker, args=(2, port, out), nprocs=2, join=True)
    g = np.load(out)
    absorbing, reflecting, smp_all, x = _rws_problem()
    smp, single = _rws_solve(smp_all, np.arange(len(smp_all), dtype=np.int64), x)
    assert np.array_equal(g["counts"], smp.group_counts())
    # the same walks (a stream per sample), summed in another order
    scale = np.abs(single.totals).max(axis=0) + 1e-6
    assert (np.abs(g["totals"] - single.totals)/scale).max() < 1e-4
