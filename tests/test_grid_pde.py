"""A PDE posed through sampled data (zb_grid_view): the way the reference's applications fill zombie::PDE's callbacks
(demo/model_problem.h:229-266 reads source, Dirichlet and Robin values from textures; include/zombie/core/pde.h:31-44).
CPU part: the shared sampler against numpy, and the oracle solving the same problem through grids and through the
closed-form descriptor. GPU part: the device against the oracle and the exact solution on problems described ONLY by grids,
including position-dependent Robin data and coefficients that no closed-form descriptor of this library can express."""
import math

import numpy as np
import pytest

from tests import cases
from zombie_b200 import abi

LO, HI = (-1.3, -1.3, -1.3), (1.3, 1.3, 1.3)


def quadratic_grids(n=65):
    u = abi.HostGrid.sample(lambda p: cases.quadratic_exact(p), LO, HI, (n, n, n))
    f = abi.HostGrid.sample(lambda p: np.full(len(p), -1.5, np.float32), LO, HI, (3, 3, 3))   # source = sigma u - lap u = -1.5
    return u, f


def grid_only_pde(**grids):
    d = abi.PdeDesc()
    d.mode = abi.PDE_CONSTANT          # every callback the problem uses is backed by a grid; the constants stay 0
    d.areRobinCoeffsNonnegative = 1
    d.areRobinConditionsPureNeumann = 0 if "robinCoeff" in grids else 1
    return d.set_grids(**grids)


def test_sampler_matches_numpy_interpolation(ref):
    """zb_grid_sample through the oracle's Dirichlet callback: on-boundary sample points report pde.dirichlet(x) verbatim
    (walk_on_spheres.h:331-345), so one solve reads the grid at arbitrary positions"""
    rng = np.random.RandomState(3)
    vals = rng.normal(size=(5, 6, 7)).astype(np.float32)           # [nz, ny, nx]
    x = rng.uniform(-1.2, 1.2, size=(200, 3)).astype(np.float32)
    (pos, tri), _, _ = cases.mesh3d(2)
    rs = ref.RefScene(3, absorbing=(pos, tri))
    for interp in (abi.GRID_NEAREST, abi.GRID_LINEAR):
        g = abi.HostGrid(vals, LO[:3], HI[:3], interp)
        pts = abi.HostSamplePoints(x, np.zeros(len(x), np.float32), None, nWalks=1, estimationQuantity=abi.SOLUTION,
                                   type=abi.ON_ABSORBING_BOUNDARY)
        st = abi.HostStatistics(len(x), 3)
        rs.solve(abi.WALK_ON_SPHERES, grid_only_pde(dirichlet=g), abi.WalkSettings.default(), pts, st, threads=1)
        t = (x.astype(np.float64) - np.array(LO))/(np.array(HI) - np.array(LO))*(np.array([7, 6, 5]) - 1)
        if interp == abi.GRID_NEAREST:
            i = np.clip(np.floor(t + 0.5).astype(int), 0, np.array([6, 5, 4]))
            want = vals[i[:, 2], i[:, 1], i[:, 0]]
            assert np.mean(st.solutionMean == want) > 0.98     # a sample within rounding of a cell face may fall either way
        else:
            i0 = np.clip(np.floor(t).astype(int), 0, np.array([5, 4, 3]))
            w = t - i0
            want = np.zeros(len(x))
            for c in range(8):
                b = np.array([(c >> k) & 1 for k in range(3)])
                wt = np.prod(np.where(b, w, 1 - w), axis=1)
                want += wt*vals[i0[:, 2] + b[2], i0[:, 1] + b[1], i0[:, 0] + b[0]]
            assert np.abs(st.solutionMean - want).max() < 1e-5
    rs.close()


def test_oracle_grid_problem_equals_closed_form_problem(ref):
    """same seeds, same walks: the only difference between the two solves is interpolation error of the data"""
    (pos, tri), _, _ = cases.mesh3d(3)
    rs = ref.RefScene(3, absorbing=(pos, tri))
    x = cases.query_points(64, 3, 0.7, seed=5)
    _, da, _, _ = rs.closest_point(abi.ABSORBING, x)
    u, f = quadratic_grids()
    out = []
    for pde in (cases.quadratic_pde(), grid_only_pde(source=f, dirichlet=u)):
        pts = abi.HostSamplePoints(x, da, None, nWalks=32, estimationQuantity=abi.SOLUTION, seed=11)
        st = abi.HostStatistics(len(x), 3)
        rs.solve(abi.WALK_ON_SPHERES, pde, abi.WalkSettings.default(), pts, st, threads=1)
        out.append(st)
    assert np.array_equal(out[0].totalWalkLength, out[1].totalWalkLength)
    assert np.abs(out[0].solutionMean - out[1].solutionMean).max() < 5e-4   # trilinear error of this quadratic on a 65^3 lattice: cell^2/8 * sum |u_kk| = 7e-4 at most
    rs.close()


@pytest.mark.gpu
def test_gpu_problem_posed_only_through_grids(ctx, ref):
    from zombie_b200 import native
    (pos, tri), _, _ = cases.mesh3d(3)
    gs, rs = native.Scene(ctx, 3, absorbing=(pos, tri)), ref.RefScene(3, absorbing=(pos, tri))
    x = cases.query_points(2048, 3, 0.7, seed=5)
    _, da, _, _ = gs.closest_point(abi.ABSORBING, x)
    u, f = quadratic_grids()
    runs = {}
    for name, solver, pde, seed in (("grid", gs, grid_only_pde(source=f, dirichlet=u), 11), ("closed", gs, cases.quadratic_pde(), 11),
                                    ("ref", rs, grid_only_pde(source=f, dirichlet=u), 12)):
        pts = abi.HostSamplePoints(x, da, None, nWalks=128, estimationQuantity=abi.SOLUTION, seed=seed)
        st = abi.HostStatistics(len(x), 3)
        solver.solve(abi.WALK_ON_SPHERES, pde, abi.WalkSettings.default(), pts, st)
        runs[name] = st
    # same seeds on the device: the grid run differs from the closed-form run by interpolation error only
    assert np.array_equal(runs["grid"].totalWalkLength, runs["closed"].totalWalkLength)
    assert np.abs(runs["grid"].solutionMean - runs["closed"].solutionMean).max() < 5e-4
    z = cases.z_scores(runs["grid"], runs["ref"])
    assert np.mean(np.abs(z) < 3.0) >= 0.99 and abs(z.mean()) < 0.1, (np.mean(np.abs(z) < 3.0), z.mean())
    se = np.sqrt(runs["grid"].solution_variance()/128)
    assert np.mean(np.abs(runs["grid"].solutionMean - cases.quadratic_exact(x)) < 3.5*se + 1e-3) > 0.98
    gs.close(); rs.close()


@pytest.mark.gpu
def test_gpu_position_dependent_robin_data_matches_reference(ctx, ref):
    """Robin values h(x) and coefficients mu(x) that vary over the boundary (mu linear in x: trilinear sampling is exact;
    per-triangle coefficient bounds = the extremes over its vertices), screened Poisson source from a grid: no closed-form
    descriptor exists for this problem, the oracle is the only reference."""
    from zombie_b200 import native
    _, absorbing, reflecting = cases.mesh3d(4)
    rp, ri = reflecting
    mu_at = lambda p: 1.0 + 0.5*p[:, 0]
    mu_v = mu_at(rp.astype(np.float64))
    mn, mx = mu_v[ri].min(axis=1).astype(np.float32), mu_v[ri].max(axis=1).astype(np.float32)
    kw = dict(absorbing=absorbing, reflecting=reflecting, reflecting_kind=abi.REFLECTING_ROBIN, min_robin=mn, max_robin=mx)
    gs, rs = native.Scene(ctx, 3, **kw), ref.RefScene(3, **kw)
    n = 49
    grids = dict(source=abi.HostGrid.sample(lambda p: np.exp(-p[:, 1])*(1.0 + p[:, 2]**2), LO, HI, (n, n, n)),
                 dirichlet=abi.HostGrid.sample(lambda p: np.cos(2.0*p[:, 0]) + p[:, 1], LO, HI, (n, n, n)),
                 robin=abi.HostGrid.sample(lambda p: 0.5*np.sin(3.0*p[:, 0]) + p[:, 1]*p[:, 2], LO, HI, (n, n, n)),
                 robinCoeff=abi.HostGrid.sample(mu_at, LO, HI, (2, 2, 2)))
    x = cases.query_points(2048, 3, 0.75, seed=21)
    _, da, _, _ = gs.closest_point(abi.ABSORBING, x)
    _, dr, _, _ = gs.closest_point(abi.REFLECTING, x)
    res = []
    for solver, seed in ((gs, 1), (rs, 2)):
        pde = grid_only_pde(**grids)
        pde.absorptionCoeff = 1.0
        pts = abi.HostSamplePoints(x, da, dr, nWalks=128, estimationQuantity=abi.SOLUTION, seed=seed)
        st = abi.HostStatistics(len(x), 3)
        out = solver.solve(abi.WALK_ON_STARS, pde, abi.WalkSettings.default(), pts, st)
        res.append((st, out if np.isscalar(out) else out[0]))
    z = cases.z_scores(res[0][0], res[1][0])
    assert np.mean(np.abs(z) < 3.0) >= 0.99 and abs(z.mean()) < 0.1 and 0.9 < z.std() < 1.12, (np.mean(np.abs(z) < 3.0), z.mean(), z.std())
    assert abs(res[0][1]/float(res[1][1]) - 1.0) < 0.03
    assert np.std(res[0][0].solutionMean) > 0.05      # the data is not trivial
    gs.close(); rs.close()
