"""Vector-valued solves (zb_solve_channels): the reference's T = Array<float, C> (python/zombie_bindings.cpp:37-44,
include/zombie/core/pde.h:31-44 — source, dirichlet and robin return T; absorptionCoeff and robinCoeff stay scalar). In the
reference the components of T share one walk and enter it only linearly, so component c of a vector-valued solve IS the scalar
solve of channel c's data on the same random numbers. That is the parity statement tested here: every channel of one
zb_solve_channels call against a scalar zb_solve of that channel's PDE with the same generators (the scalar path itself is
pinned to the compiled reference by tests/test_gpu_solve.py) — identical walk counts and lengths, estimates equal up to the
contraction of multiply-adds — plus the exact solutions, progressive refinement and the argument checks."""
import ctypes as C

import numpy as np
import pytest

from tests import cases
from zombie_b200 import abi, native, workloads

pytestmark = pytest.mark.gpu


def _like(first, other):
    """a copy of `other` carrying what steers the walk from `first` (the channels share it)"""
    d = abi.PdeDesc()
    C.memmove(C.byref(d), C.byref(other), C.sizeof(abi.PdeDesc))
    for f in ("absorptionCoeff", "isSourceConstant", "areRobinConditionsPureNeumann", "areRobinCoeffsNonnegative", "robinCoeff", "kelvin"):
        setattr(d, f, getattr(first, f))
    return d


def _constant(first, source, dirichlet, robin):
    d = _like(first, abi.PdeDesc())
    d.mode = abi.PDE_CONSTANT
    d.sourceValue, d.dirichletValue, d.robinValue = source, dirichlet, robin
    return d


def _points(scene, x, n_walks, quantity, seed=0xC0FFEE):
    _, da, _, _ = scene.closest_point(abi.ABSORBING, x)
    dr = None
    if scene.stats(abi.REFLECTING)["nodes"]:
        _, dr, _, _ = scene.closest_point(abi.REFLECTING, x)
    return abi.HostSamplePoints(x, da, dr, nWalks=n_walks, estimationQuantity=quantity, seed=seed)


def _assert_same(a, b, what):
    assert np.array_equal(a.nSolutionEstimates, b.nSolutionEstimates), what
    assert np.array_equal(a.nGradientEstimates, b.nGradientEstimates), what
    assert np.array_equal(a.totalWalkLength, b.totalWalkLength), what
    assert np.array_equal(a.totalSplits, b.totalSplits), what
    scale = max(1.0, float(np.abs(b.solutionMean).max()))
    assert np.abs(a.solutionMean - b.solutionMean).max() <= 2e-5*scale, what
    assert np.abs(a.totalFirstSourceContribution - b.totalFirstSourceContribution).max() <= 1e-4*max(1.0, float(np.abs(b.totalFirstSourceContribution).max())), what
    gscale = max(1.0, float(np.abs(b.gradientMean).max()))
    assert np.abs(a.gradientMean - b.gradientMean).max() <= 1e-4*gscale, what
    assert np.abs(a.solutionM2 - b.solutionM2).max() <= 1e-3*max(1.0, float(b.solutionM2.max())), what
    assert np.abs(a.totalDerivativeContribution - b.totalDerivativeContribution).max() <= 1e-3*max(1.0, float(np.abs(b.totalDerivativeContribution).max())), what


def _run(scene, estimator, pdes, settings, x, n_walks, quantity, repeats=1):
    dim = x.shape[1]
    pts = _points(scene, x, n_walks, quantity)
    multi = [abi.HostStatistics(len(x), dim) for _ in pdes]
    steps = [scene.solve_channels(estimator, pdes, settings, pts, multi) for _ in range(repeats)]
    single = []
    for pde in pdes:
        p = _points(scene, x, n_walks, quantity)
        st = abi.HostStatistics(len(x), dim)
        s = [scene.solve(estimator, pde, settings, p, st) for _ in range(repeats)]
        assert s == steps
        assert np.array_equal(p.rngState, pts.rngState)
        single.append(st)
    return multi, single


def test_three_channels_walk_on_stars_robin_gradients(ctx):
    """C3's problem (3D screened Poisson, Robin + Dirichlet, solution and gradient) with three value channels"""
    w = workloads.config3(subdivisions=4, n_points=1024, n_walks=32)
    scene = native.Scene(ctx, 3, **w.scene_kwargs())
    k2 = np.array([0.0, 0.6, -0.8])*np.sqrt(w.pde.absorptionCoeff)
    pdes = [w.pde,
            _like(w.pde, abi.PdeDesc.manufactured(abi.SOL_EXP_DIR, list(k2) + [0.5], absorptionCoeff=w.pde.absorptionCoeff, robinCoeff=w.pde.robinCoeff)),
            _constant(w.pde, 0.7, -1.3, 0.4)]
    multi, single = _run(scene, w.estimator, pdes, w.settings, w.points, w.n_walks, w.quantity)
    for c in range(3):
        _assert_same(multi[c], single[c], "channel %d" % c)
    # channel 0 and 1 against their exact solutions
    for c, exact in ((0, w.exact(w.points)), (1, 0.5*np.exp(w.points.astype(np.float64) @ k2))):
        st = multi[c]
        se = np.sqrt(st.solution_variance()/np.maximum(1, st.nSolutionEstimates))
        assert np.mean(np.abs(st.solutionMean - exact) < 3.0*se + 2e-3) >= 0.97
    assert not np.allclose(multi[0].solutionMean, multi[1].solutionMean)
    scene.close()


def test_channels_walk_on_spheres_2d_with_source_and_progressive_refinement(ctx):
    """2D Poisson, Walk on Spheres, non-constant source (sampled in the ball), two solves in a row: the statistics of every channel
    accumulate like channel 0's"""
    (pos, seg), _, _ = cases.mesh2d(1024)
    scene = native.Scene(ctx, 2, absorbing=(pos, seg))
    first = abi.PdeDesc.manufactured(abi.SOL_QUADRATIC, [1.0, -0.5, 0.0, 0.3])
    pdes = [first, _like(first, abi.PdeDesc.manufactured(abi.SOL_CUBIC_2D, [0.0]*4)),
            _like(first, abi.PdeDesc.manufactured(abi.SOL_QUADRATIC, [0.25, 2.0, 0.0, -1.0]))]
    x = cases.query_points(800, 2, 0.6, seed=9)
    settings = abi.WalkSettings.default()
    multi, single = _run(scene, abi.WALK_ON_SPHERES, pdes, settings, x, 48, abi.SOLUTION_AND_GRADIENT, repeats=2)
    for c in range(3):
        _assert_same(multi[c], single[c], "channel %d" % c)
        assert (multi[c].nSolutionEstimates == 96).all()
    exact1 = x[:, 0].astype(np.float64)**3 - 3.0*x[:, 0]*x[:, 1].astype(np.float64)**2
    se = np.sqrt(multi[1].solution_variance()/96.0)
    assert np.mean(np.abs(multi[1].solutionMean - exact1) < 3.0*se + 2e-3) >= 0.97
    scene.close()


def test_channels_neumann_weight_window_and_boundary_points(ctx):
    """Walk on Stars with a Neumann boundary and the demo's weight window (splits and roulette: walk states that end early or
    multiply), sample points ON the absorbing boundary (each channel reports its own Dirichlet value) and a skipped point"""
    whole, absorbing, reflecting = cases.mesh3d(3)
    scene = native.Scene(ctx, 3, **cases.scene_kwargs(3, abi.REFLECTING_NEUMANN, absorbing, reflecting))
    first = cases.quadratic_pde()
    pdes = [first, _like(first, abi.PdeDesc.manufactured(abi.SOL_QUADRATIC, [0.5, 0.5, -1.0, 2.0])), _constant(first, 1.0, 0.0, 0.0)]
    settings = abi.WalkSettings.default()
    settings.russianRouletteThreshold, settings.splittingThreshold = 0.99, 1.5
    x = cases.query_points(600, 3, 0.8, seed=13)
    dim = 3
    pts = _points(scene, x, 24, abi.SOLUTION)
    on_boundary = np.arange(0, 600, 7)
    verts = absorbing[0][absorbing[1][:, 0]]
    pts.pt[on_boundary] = verts[:len(on_boundary)]
    pts.type[on_boundary] = abi.ON_ABSORBING_BOUNDARY
    pts.distToAbsorbingBoundary[on_boundary] = 0.0
    pts.estimationQuantity[5] = abi.SKIP
    import copy
    multi = [abi.HostStatistics(len(x), dim) for _ in pdes]
    base = copy.deepcopy(pts)
    scene.solve_channels(abi.WALK_ON_STARS, pdes, settings, pts, multi)
    for c, pde in enumerate(pdes):
        p = copy.deepcopy(base)
        st = abi.HostStatistics(len(x), dim)
        scene.solve(abi.WALK_ON_STARS, pde, settings, p, st)
        _assert_same(multi[c], st, "channel %d" % c)
    assert (multi[1].nSolutionEstimates[on_boundary] == 1).all() and multi[1].nSolutionEstimates[5] == 0
    assert not np.allclose(multi[0].solutionMean[on_boundary], multi[1].solutionMean[on_boundary])   # each channel reports ITS Dirichlet value
    scene.close()


def test_channel_argument_checks(ctx):
    (pos, seg), _, _ = cases.mesh2d(256)
    scene = native.Scene(ctx, 2, absorbing=(pos, seg))
    first = abi.PdeDesc.manufactured(abi.SOL_CUBIC_2D, [0.0]*4)
    other = abi.PdeDesc.manufactured(abi.SOL_CUBIC_2D, [0.0]*4, absorptionCoeff=1.0)
    x = cases.query_points(16, 2, 0.5, seed=1)
    pts = _points(scene, x, 4, abi.SOLUTION)
    stats = [abi.HostStatistics(16, 2), abi.HostStatistics(16, 2)]
    with pytest.raises(native.ZombieError, match="share"):
        scene.solve_channels(abi.WALK_ON_SPHERES, [first, other], abi.WalkSettings.default(), pts, stats)
    with pytest.raises(native.ZombieError):
        scene.solve_channels(abi.WALK_ON_SPHERES, [first], abi.WalkSettings.default(), pts, stats)
    # one channel is the scalar solve
    st1, st2 = abi.HostStatistics(16, 2), abi.HostStatistics(16, 2)
    scene.solve_channels(abi.WALK_ON_SPHERES, [first], abi.WalkSettings.default(), _points(scene, x, 4, abi.SOLUTION), [st1])
    scene.solve(abi.WALK_ON_SPHERES, first, abi.WalkSettings.default(), _points(scene, x, 4, abi.SOLUTION), st2)
    assert np.array_equal(st1.solutionMean, st2.solutionMean)
    scene.close()


def test_device_resident_channels_match_host_buffers(ctx):
    """zb_solve_channels_dev on torch tensors == zb_solve_channels on host arrays (same seeds -> identical statistics, every channel)"""
    import torch
    w = workloads.config3(subdivisions=3, n_points=256, n_walks=16)
    scene = native.Scene(ctx, 3, **w.scene_kwargs())
    pdes = [w.pde, _constant(w.pde, 0.7, -1.3, 0.4)]
    pts = _points(scene, w.points, w.n_walks, w.quantity)
    host = [abi.HostStatistics(pts.n, 3) for _ in pdes]
    hp = _points(scene, w.points, w.n_walks, w.quantity)
    steps_host = scene.solve_channels(w.estimator, pdes, w.settings, hp, host)
    dev = torch.device("cuda", 0)
    keys_p = ("pt", "normal", "directionForDerivative", "type", "estimationQuantity", "distToAbsorbingBoundary",
              "distToReflectingBoundary", "estimateBoundaryNormalAligned", "nWalks", "rngState", "rngInc", "firstSphereRadius")
    keys_s = ("solutionMean", "solutionM2", "gradientMean", "gradientM2", "totalFirstSourceContribution",
              "totalDerivativeContribution", "nSolutionEstimates", "nGradientEstimates", "totalWalkLength", "totalSplits")
    tp = {}
    for k in keys_p:
        a = getattr(pts, k)
        tp[k] = torch.from_numpy(a.view(np.int64) if a.dtype == np.uint64 else a).to(dev)
    ts = [{k: torch.from_numpy(getattr(abi.HostStatistics(pts.n, 3), k)).to(dev) for k in keys_s} for _ in pdes]
    steps = torch.zeros(1, dtype=torch.int64, device=dev)
    torch.cuda.synchronize()
    ps = abi.SamplePoints(pts.n, *[tp[k].data_ptr() for k in keys_p])
    ss = [abi.Statistics(*[t[k].data_ptr() for k in keys_s]) for t in ts]
    scene.solve_channels_dev(w.estimator, pdes, w.settings, ps, ss, steps.data_ptr())
    ctx.synchronize()
    assert int(steps.item()) == steps_host
    for c in range(2):
        for k in ("solutionMean", "solutionM2", "gradientMean", "totalWalkLength", "nSolutionEstimates", "totalDerivativeContribution"):
            assert np.array_equal(ts[c][k].cpu().numpy(), getattr(host[c], k)), (c, k)
    assert np.array_equal(tp["rngState"].cpu().numpy().view(np.uint64), hp.rngState)
    scene.close()
