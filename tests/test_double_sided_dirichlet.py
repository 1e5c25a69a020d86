"""Double-sided absorbing boundaries with different Dirichlet data on their two sides (the reference's diffusion-curve demo,
demo/model_problems/nautilus/wos.json: solveDoubleSided with absorbingBoundaryValue / absorbingBoundaryNormalAlignedValue):
a walk that ends on the side the closest primitive's normal points to reads pde.dirichlet(x, true), on the other side
pde.dirichlet(x, false) (walk_on_spheres.h:300-308, walk_on_stars.h:499-506; fcpw Interaction::signedDistance, interaction.h:32-34).
zb_pde_desc carries the second side as dirichletAlignedGrid / dirichletAlignedValue.

CPU: the compiled reference on the test problem pins the convention. GPU: the device against it."""
import ctypes as C

import numpy as np
import pytest

from tests import cases
from zombie_b200 import abi, procedural as P, zombie

LO, HI = np.array([-2.0, -2.0], np.float32), np.array([2.0, 2.0], np.float32)


def problem(n=600, seed=0):
    """closed star polygon (outward normals) as a double-sided absorbing curve inside the box [-2, 2]^2, not watertight"""
    pos, seg = P.star_polygon(512)
    x = np.random.RandomState(seed).uniform(-1.6, 1.6, size=(n, 2)).astype(np.float32)
    r, th = np.linalg.norm(x, axis=1), np.arctan2(x[:, 1], x[:, 0])
    rb = 1.0 + 0.25*np.sin(5.0*th)
    keep = np.abs(r - rb) > 0.02
    return pos, seg, x[keep], (r < rb)[keep]


def scene_kwargs(pos, seg):
    return dict(absorbing=(pos, seg), solve_double_sided=True, watertight=False, domain_min=LO, domain_max=HI)


def constant_pde(inner=1.0, outer=3.0):
    d = abi.PdeDesc()
    d.mode = abi.PDE_CONSTANT
    d.areRobinConditionsPureNeumann, d.areRobinCoeffsNonnegative, d.isSourceConstant = 1, 1, 1
    d.dirichletValue, d.dirichletAlignedValue, d.hasDirichletAlignedValue = inner, outer, 1
    return d


def grid_pde():
    d = abi.PdeDesc()
    d.mode = abi.PDE_CONSTANT
    d.areRobinConditionsPureNeumann, d.areRobinCoeffsNonnegative, d.isSourceConstant = 1, 1, 1
    a = abi.HostGrid.sample(lambda p: 0.5 + 0.3*p[:, 0], (-2.0, -2.0, 0.0), (2.0, 2.0, 0.0), (65, 65, 1))
    b = abi.HostGrid.sample(lambda p: -1.0 + 0.2*p[:, 1], (-2.0, -2.0, 0.0), (2.0, 2.0, 0.0), (65, 65, 1))
    return d.set_grids(dirichlet=a, dirichletAligned=b)


def solve(scene, pde, x, n_walks=32, seed=0xC0FFEE):
    _, da, _, _ = scene.closest_point(abi.ABSORBING, x)
    pts = abi.HostSamplePoints(x, da, None, nWalks=n_walks, estimationQuantity=abi.SOLUTION, seed=seed)
    st = abi.HostStatistics(len(x), 2)
    scene.solve(abi.WALK_ON_SPHERES, pde, abi.WalkSettings.default(1e-3, 1e-3, 1024, True), pts, st)
    return st


def test_reference_convention(ref):
    """inside the curve walks end on the side opposite to the (outward) normals and read dirichlet(x, false); outside they read
    dirichlet(x, true) — except the few that end at a vertex whose closest segment faces away (the sign is that of ONE primitive)"""
    pos, seg, x, inside = problem()
    st = solve(ref.RefScene(2, **scene_kwargs(pos, seg)), constant_pde(), x)
    ok = st.nSolutionEstimates > 0
    ins = st.solutionMean[inside]
    assert ok[inside].all() and (ins < 1.5).all() and (ins == 1.0).mean() > 0.8 and ins.mean() < 1.05
    out = st.solutionMean[~inside & ok]
    assert (out > 2.5).all() and (out == 3.0).mean() > 0.8 and out.mean() > 2.95


def test_surface_passes_the_aligned_dirichlet_grid():
    data = np.arange(12, dtype=np.float32)
    cb = zombie.Utils.get_dense_grid_dirichlet_callback(data, 2.0*data, [4, 3], np.zeros(2), np.array([4.0, 3.0]), dim=2)
    pde = zombie.Core.PDE(dim=2)
    pde.dirichlet = cb
    d = pde.descriptor()
    assert d.dirichletGrid.data and d.dirichletAlignedGrid.data
    assert np.array_equal(d.host_grids["dirichletAlignedGrid"].values, 2.0*d.host_grids["dirichletGrid"].values)
    assert C.sizeof(abi.PdeDesc) == abi.PdeDesc.dirichletAlignedGrid.offset + C.sizeof(abi.GridView) + 8


@pytest.mark.gpu
def test_device_reads_the_side_the_walk_ends_on(ctx, ref):
    from zombie_b200 import native
    pos, seg, x, inside = problem()
    gs, rs = native.Scene(ctx, 2, **scene_kwargs(pos, seg)), ref.RefScene(2, **scene_kwargs(pos, seg))
    g, r = solve(gs, constant_pde(), x), solve(rs, constant_pde(), x)
    ok = g.nSolutionEstimates > 0
    ins, rins = g.solutionMean[inside], r.solutionMean[inside]
    assert ok[inside].all() and (ins < 1.5).all() and (ins == 1.0).mean() > 0.8
    out = g.solutionMean[~inside & ok]
    rout = r.solutionMean[~inside & (r.nSolutionEstimates > 0)]
    assert (out > 2.5).all() and (out == 3.0).mean() > 0.8
    # which primitive answers at a vertex follows each tree: the same rate of walks on the other side, not the same walks
    assert abs(out.mean() - rout.mean()) < 0.02 and abs(ins.mean() - rins.mean()) < 0.02
    # the same two sides with swapped values: nothing but the data changes
    g2 = solve(gs, constant_pde(3.0, 1.0), x)
    assert np.allclose(g2.solutionMean[ok], 4.0 - g.solutionMean[ok], atol=1e-6)
    # single-sided solves never read the second side
    one = abi.WalkSettings.default(1e-3, 1e-3, 1024, False)
    _, da, _, _ = gs.closest_point(abi.ABSORBING, x)
    pts = abi.HostSamplePoints(x, da, None, nWalks=8, estimationQuantity=abi.SOLUTION)
    st = abi.HostStatistics(len(x), 2)
    gs.solve(abi.WALK_ON_SPHERES, constant_pde(), one, pts, st)
    assert (st.solutionMean[st.nSolutionEstimates > 0] == 1.0).all()
    gs.close(); rs.close()


@pytest.mark.gpu
def test_device_matches_reference_on_two_sided_grid_data(ctx, ref):
    from zombie_b200 import native
    pos, seg, x, inside = problem(n=900, seed=3)
    gs, rs = native.Scene(ctx, 2, **scene_kwargs(pos, seg)), ref.RefScene(2, **scene_kwargs(pos, seg))
    g, r = solve(gs, grid_pde(), x, n_walks=64, seed=1), solve(rs, grid_pde(), x, n_walks=64, seed=2)
    ok = (g.nSolutionEstimates > 8) & (r.nSolutionEstimates > 8)
    z = cases.z_scores(g, r)[ok]
    assert np.mean(np.abs(z) < 3.0) >= 0.99 and abs(float(np.mean(z))) < 0.15
    # inside the curve the solution is the harmonic extension of 0.5 + 0.3 x, i.e. that function itself
    err = g.solutionMean[inside] - (0.5 + 0.3*x[inside, 0])
    se = np.sqrt(g.solution_variance()[inside]/np.maximum(1, g.nSolutionEstimates[inside]))
    assert np.mean(np.abs(err) < 3.0*se + 5e-3) > 0.98
    # sample points ON the curve report the side their flag names (walk_on_spheres.h:331-337)
    p = np.ascontiguousarray(pos[seg[::16, 0]], np.float32)
    pts = abi.HostSamplePoints(p, np.zeros(len(p), np.float32), None, nWalks=1, estimationQuantity=abi.SOLUTION, type=abi.ON_ABSORBING_BOUNDARY)
    pts.estimateBoundaryNormalAligned[::2] = 1
    st = abi.HostStatistics(len(p), 2)
    gs.solve(abi.WALK_ON_SPHERES, grid_pde(), abi.WalkSettings.default(1e-3, 1e-3, 1024, True), pts, st)
    want = np.where(pts.estimateBoundaryNormalAligned != 0, -1.0 + 0.2*p[:, 1], 0.5 + 0.3*p[:, 0])
    assert np.abs(st.solutionMean - want).max() < 2e-3
    gs.close(); rs.close()


@pytest.mark.gpu
def test_channels_carry_their_own_two_sides(ctx):
    from zombie_b200 import native
    pos, seg, x, inside = problem(n=300, seed=5)
    gs = native.Scene(ctx, 2, **scene_kwargs(pos, seg))
    pdes = [constant_pde(1.0, 3.0), constant_pde(-2.0, 7.0)]
    _, da, _, _ = gs.closest_point(abi.ABSORBING, x)
    pts = abi.HostSamplePoints(x, da, None, nWalks=16, estimationQuantity=abi.SOLUTION)
    stats = [abi.HostStatistics(len(x), 2) for _ in pdes]
    gs.solve_channels(abi.WALK_ON_SPHERES, pdes, abi.WalkSettings.default(1e-3, 1e-3, 1024, True), pts, stats)
    ok = stats[0].nSolutionEstimates > 0
    assert (stats[0].solutionMean[inside] == 1.0).mean() > 0.8 and (stats[1].solutionMean[inside] == -2.0).mean() > 0.8
    # both channels saw the same sides: channel 1 = -2 + 9 * (fraction of walks that ended on the aligned side)
    frac = (stats[0].solutionMean[ok] - 1.0)/2.0
    assert np.allclose(stats[1].solutionMean[ok], -2.0 + 9.0*frac, atol=1e-5)
    gs.close()
