"""The reference's 2D demo model problem (demo/model_problems/engine/wost.json: Walk on Stars, 256 x 256 grid, 128 walks per point,
weight window 0.99 / 1.5, texture-backed PDE data) run through zombie_b200's Python surface with the call sequence of the
reference's demo (demo/demo.cpp:9-46, 560-620; demo/model_problem.h:129-300; demo/grid.h:22-85), and compared with the solution
image the reference committed for it (demo/model_problems/engine/solutions/wost.pfm) — the only end-to-end artefact of the
estimator the reference holds. Inputs and the reference image travel as tests/golden/demo_engine.npz
(tests/golden/make_demo_fixture.py). Both images are 128-walk Monte Carlo estimates, so the gate is statistical per pixel.

Also: the same problem with four value channels (the demo's `channels: 4`, T = Array4: the colour textures' R, G, B and a zero
channel) — by linearity 0.299 R + 0.587 G + 0.114 B of that solve is the scalar (luminance) solve, walk for walk."""
import json
import os

import numpy as np
import pytest

from tests.conftest import GOLDEN
from zombie_b200 import pfm, zombie

FIXTURE = os.path.join(GOLDEN, "demo_engine.npz")


@pytest.fixture(scope="module")
def engine():
    return np.load(FIXTURE)


def test_pfm_round_trip(tmp_path):
    rng = np.random.RandomState(1)
    for shape in ((5, 7, 3), (4, 6, 1)):
        img = rng.uniform(-2.0, 30.0, size=shape).astype(np.float32)
        p = str(tmp_path / "a.pfm")
        pfm.write_pfm(p, img)
        # Image::write stores the bottom row first (image.h:283-291); Image::read keeps the file order (image.h:222-240)
        assert np.array_equal(pfm.read_pfm(p), img[::-1])
    assert pfm.luminance(np.array([[[1.0, 1.0, 1.0]]], np.float32))[0, 0] == pytest.approx(1.0, abs=1e-6)


def texture_lookup(buffer, uv):
    """Image<C>::get(uv, flipY = true) (demo/image.h:103-110) on an image in buffer order"""
    h, w = buffer.shape[:2]
    i = np.clip((uv[:, 1]*h).astype(int), 0, h - 1)
    j = np.clip((uv[:, 0]*w).astype(int), 0, w - 1)
    return buffer[h - 1 - i, j]


def dense_grid_data(buffer):
    """the demo's textures as DenseGrid data ([nx, ny(, channels)], y fastest, cell (i, j) centred at uv = ((i + .5)/w, (j + .5)/h)):
    Image::get(uv) reads row h - 1 - int(uv.y h), column int(uv.x w)"""
    return np.ascontiguousarray(np.swapaxes(buffer[::-1], 0, 1))


def test_fixture_and_texture_convention(engine, tmp_path):
    obj = str(tmp_path / "geometry.obj")
    open(obj, "wb").write(engine["obj"].tobytes())
    positions, indices = zombie.FloatNList(dim=2), zombie.IntNList(dim=2)
    zombie.Utils.load_boundary_mesh(obj, positions, indices, dim=2)
    assert len(positions) == 647 and len(indices) == 647
    assert engine["ref_wost"].shape == (256, 256) and engine["absorbing"].shape == (512, 512, 3)
    # the dense-grid callback of the surface reads the texture exactly like the demo's nearest-pixel lookup
    lum = pfm.luminance(engine["absorbing"])
    lo, hi = np.array([-1.0, -1.0]), np.array([1.0, 1.0])
    cb = zombie.Utils.get_dense_grid_dirichlet_callback(dense_grid_data(lum).ravel(), [512, 512], lo, hi, dim=2)
    rng = np.random.RandomState(3)
    x = rng.uniform(-1.0, 1.0, size=(2000, 2))
    want = texture_lookup(lum, (x - lo)/2.0)
    got = np.array([cb(p) for p in x])
    assert np.mean(got == want) > 0.995     # positions within rounding of a pixel edge may fall either way


def setup_problem(engine, tmp_path, channels, config="wost"):
    """ModelProblem (model_problem.h:129-207) + the evaluation grid and its distance info (grid.h:22-40, demo.cpp:9-24)"""
    cfg = json.loads(engine["config_" + config].tobytes().decode())
    solver_cfg, out_cfg, mp = cfg["solver"], cfg["output"], cfg["modelProblem"]
    dim = 2
    obj = str(tmp_path / "geometry.obj")
    open(obj, "wb").write(engine["obj"].tobytes())
    # ModelProblem::loadOBJ (model_problem.h:209-215)
    positions, indices = zombie.FloatNList(dim=dim), zombie.IntNList(dim=dim)
    zombie.Utils.load_boundary_mesh(obj, positions, indices, dim=dim)
    zombie.Utils.normalize(positions, dim=dim)
    zombie.Utils.flip_orientation(indices, dim=dim)
    bbox = zombie.Utils.compute_bounding_box(positions, True, 1.0, dim=dim)

    # ModelProblem::setupPDE (:230-276): T = float reads the luminance of the textures, T = Array4 their R, G, B and a zero channel
    def texture(name):
        rgb = engine[name]
        if channels == 1:
            return dense_grid_data(pfm.luminance(rgb)).ravel()
        return dense_grid_data(np.concatenate([rgb, np.zeros_like(rgb[..., :1])], axis=-1)).reshape(-1, 4)

    shape = [512, 512]
    pde = zombie.Core.PDE(dim=dim, channels=channels)
    pde.source = zombie.Utils.get_dense_grid_source_callback(texture("source"), shape, bbox[0], bbox[1], dim=dim, channels=channels)
    pde.dirichlet = zombie.Utils.get_dense_grid_dirichlet_callback(texture("absorbing"), shape, bbox[0], bbox[1], dim=dim, channels=channels)
    pde.robin = zombie.Utils.get_dense_grid_robin_callback(texture("reflecting"), shape, bbox[0], bbox[1], dim=dim, channels=channels)
    pde.robin_coeff = zombie.Core.get_constant_robin_coefficient_callback(mp["robinCoeff"], dim=dim)
    pde.has_reflecting_boundary_conditions = zombie.Utils.get_dense_grid_indicator_callback(
        dense_grid_data(pfm.luminance(engine["is_reflecting"])).ravel(), [1024, 1024], bbox[0], bbox[1], dim=dim)
    pde.absorption_coeff = mp["absorptionCoeff"]
    pde.is_source_constant = False
    pde.are_robin_conditions_pure_neumann = mp["robinCoeff"] == 0.0
    pde.are_robin_coeffs_nonnegative = mp["robinCoeff"] >= 0.0

    # partitionBoundaryMesh + setupGeometricQueries (:278-340)
    ap, ai, rp, ri = (zombie.FloatNList(dim=dim), zombie.IntNList(dim=dim), zombie.FloatNList(dim=dim), zombie.IntNList(dim=dim))
    zombie.Utils.partition_boundary_mesh(pde.has_reflecting_boundary_conditions, positions, indices, ap, ai, rp, ri, dim=dim)
    assert len(ai) > 0 and len(ri) > 0 and len(ai) + len(ri) == len(indices)
    queries = zombie.Core.GeometricQueries(mp["domainIsWatertight"], bbox[0], bbox[1], dim=dim)
    dirichlet = zombie.Utils.FcpwDirichletBoundaryHandler(dim=dim)
    dirichlet.build_acceleration_structure(ap, ai)
    zombie.Utils.populate_geometric_queries_for_dirichlet_boundary(dirichlet, queries, dim=dim)
    neumann = zombie.Utils.FcpwNeumannBoundaryHandler(dim=dim)
    neumann.build_acceleration_structure(rp, ri, zombie.Utils.get_ignore_candidate_silhouette_callback(mp["solveDoubleSided"]))
    zombie.Utils.populate_geometric_queries_for_neumann_boundary(neumann, zombie.Utils.get_branch_traversal_weight_callback(), queries, dim=dim)

    # createGridPoints (grid.h:22-40), computeDistanceInfo + createSamplePoints (demo.cpp:9-46)
    res = out_cfg["gridRes"]
    ii, jj = np.meshgrid(np.arange(res), np.arange(res), indexing="ij")            # index = i*res + j
    extent = bbox[1] - bbox[0]
    grid = np.stack([(ii.ravel()/np.float32(res))*extent[0] + bbox[0][0], (jj.ravel()/np.float32(res))*extent[1] + bbox[0][1]], axis=1).astype(np.float32)
    solve_locations = zombie.FloatNList(dim=dim)
    solve_locations.extend(list(grid))
    da, dr = zombie.FloatList(), zombie.FloatList()
    zombie.Utils.compute_dist_to_boundary(queries, solve_locations, da, dr, dim=dim)
    inside = queries.inside_domain_batch(grid).astype(bool)
    where = np.where(inside)[0]
    meshes = tuple(np.array(a, t) for a, t in ((ap, np.float32), (ai, np.int32), (rp, np.float32), (ri, np.int32)))
    return dict(cfg=cfg, pde=pde, queries=queries, bbox=bbox, grid=grid, da=np.array(da, np.float32), dr=np.array(dr, np.float32),
                where=where, res=res, meshes=meshes)


def to_image(problem, values, channels=1):
    """saveGridValues (grid.h:43-85): image(j, i) = value at grid index i*res + j, zero outside the domain and inside the boundary mask"""
    res, where = problem["res"], problem["where"]
    dist = np.minimum(problem["da"], problem["dr"])
    image = np.zeros((res, res) if channels == 1 else (res, res, channels), np.float32)
    visible = np.zeros((res, res), bool)
    keep = dist[where] >= problem["cfg"]["output"]["boundaryDistanceMask"]
    i, j = np.divmod(where[keep], res)
    image[j, i] = np.asarray(values)[keep]
    visible[j, i] = True
    return image, visible


def run_demo(engine, tmp_path, channels):
    """runWalkOnStars (demo.cpp:95-150) on the problem of engine/wost.json"""
    P = setup_problem(engine, tmp_path, channels)
    cfg, pde, queries, grid, da, dr, where, res = (P[k] for k in ("cfg", "pde", "queries", "grid", "da", "dr", "where", "res"))
    solver_cfg, out_cfg, mp = cfg["solver"], cfg["output"], cfg["modelProblem"]
    dim = 2
    zombie._SamplePoint._next_stream = 0x5EED
    sample_pts = zombie.Solvers.SamplePointList(
        [zombie.Solvers.SamplePoint(grid[k], np.zeros(dim), zombie.Solvers.SampleType.InDomain, zombie.Solvers.EstimationQuantity.Solution,
                                    1.0, da[k], dr[k], dim=dim, channels=channels) for k in where], dim=dim, channels=channels)

    # runWalkOnStars (demo.cpp:100-150): the long WalkSettings constructor with the config's values
    settings = zombie.Solvers.WalkSettings(solver_cfg["epsilonShellForAbsorbingBoundary"], solver_cfg["epsilonShellForReflectingBoundary"],
                                           1e-3, solver_cfg["russianRouletteThreshold"], solver_cfg["splittingThreshold"],
                                           solver_cfg["maxWalkLength"], 0, solver_cfg["maxWalkLength"], mp["solveDoubleSided"], True, True, False,
                                           solver_cfg["ignoreAbsorbingBoundaryContribution"], solver_cfg["ignoreReflectingBoundaryContribution"],
                                           solver_cfg["ignoreSourceContribution"], False)
    stats = zombie.Solvers.create_sample_statistics_list(len(sample_pts), dim=dim, channels=channels)
    solver = zombie.Solvers.WalkOnStars(queries, dim=dim, channels=channels)
    solver.solve(pde, settings, zombie.IntList([solver_cfg["nWalks"]]*len(sample_pts)), sample_pts, stats)

    n = len(where)
    image, visible = to_image(P, np.array([stats[k].get_estimated_solution() for k in range(n)]), channels)
    variance, _ = to_image(P, np.array([stats[k].get_estimated_solution_variance() for k in range(n)]), channels)
    return image, variance, visible, solver_cfg["nWalks"]


class TextureData:
    """host values of the demo's PDE callbacks at arrays of points (model_problem.h:230-262), for the cache bookkeeping of
    zombie_b200.bvc: Image::get((x - bMin)/maxLength) of the luminance textures"""

    def __init__(self, engine, bbox, robin_coeff):
        self.lo, self.length = np.asarray(bbox[0], np.float64), float(np.max(bbox[1] - bbox[0]))
        self.textures = {k: pfm.luminance(engine[k]) for k in ("source", "reflecting")}
        self.mu = robin_coeff

    def _get(self, name, x):
        return texture_lookup(self.textures[name], (np.asarray(x, np.float64) - self.lo)/self.length).astype(np.float32)

    def source(self, x):
        return self._get("source", x)

    def robin(self, x, n):
        return self._get("reflecting", x)

    def robin_coeff(self, x, n):
        return np.full(len(x), self.mu, np.float32)


def run_demo_bvc(engine, tmp_path):
    """runBoundaryValueCaching (demo.cpp:199-319) on the problem of engine/bvc.json, over zombie_b200.bvc"""
    from zombie_b200 import abi, bvc
    P = setup_problem(engine, tmp_path, 1, "bvc")
    cfg, pde, queries, bbox, grid, da, dr, where = (P[k] for k in ("cfg", "pde", "queries", "bbox", "grid", "da", "dr", "where"))
    sc, mp = cfg["solver"], cfg["modelProblem"]
    eps_a = sc["epsilonShellForAbsorbingBoundary"]
    settings = zombie.Solvers.WalkSettings(eps_a, sc["epsilonShellForReflectingBoundary"], 1e-3, sc["russianRouletteThreshold"],
                                           sc["splittingThreshold"], sc["maxWalkLength"], 0, sc["maxWalkLength"], mp["solveDoubleSided"], True, True,
                                           False, sc["ignoreAbsorbingBoundaryContribution"], sc["ignoreReflectingBoundaryContribution"],
                                           sc["ignoreSourceContribution"], False)
    offset_a, offset_r = 5.0*eps_a, 0.0                      # normalOffsetFor{Absorbing,Reflecting}Boundary defaults (:240-242)
    ap, ai, rp, ri = P["meshes"]
    # queries.insideBoundingDomain (:181-187), exact comparisons. The engine's outline has long segments lying ON the bounding box;
    # they pass because computeBoundingBox pads the box by epsilon (without the padding 8.8 % of the reflecting boundary drops out
    # of the sampler by an ulp and the cache misses that part of the boundary integral: -0.09 on the whole image)
    inside_box = lambda p: np.all((p >= bbox[0]) & (p <= bbox[1]), axis=1)
    sa = bvc.UniformBoundarySampler(ap, ai, inside_solve_region=inside_box, seed=21)
    sr = bvc.UniformBoundarySampler(rp, ri, inside_solve_region=inside_box, seed=22)
    sa.initialize(offset_a, mp["solveDoubleSided"])
    sr.initialize(offset_r, mp["solveDoubleSided"])
    assert sc["ignoreSourceContribution"]                     # -> domainCacheSize = 0 (:271)
    backend = bvc.NativeBackend(zombie._context(), queries.scene())
    solver = bvc.BoundaryValueCachingSolver(backend, sa, sr, None, seed=23)
    solver.generate_samples(sc["absorbingBoundaryCacheSize"], sc["reflectingBoundaryCacheSize"], 0, offset_a, offset_r)
    desc = pde.descriptor()
    solver.compute_sample_estimates(desc, settings, sc["nWalksForCachedSolutionEstimates"], sc["nWalksForCachedGradientEstimates"], abi.FLT_MAX,
                                    TextureData(engine, bbox, mp["robinCoeff"]))
    e = abi.HostBvcEvalPoints(grid[where], da[where], dr[where])
    solver.splat(desc, e, sc["radiusClampForKernels"], 0.0, abi.FLT_MAX, offset_a, offset_r)
    solver.estimate_solution_near_boundary(desc, settings, e, offset_a, offset_r, sc["nWalksForCachedSolutionEstimates"])
    image, visible = to_image(P, e.solution)
    return image, visible, solver


@pytest.mark.gpu
def test_demo_engine_walk_on_stars_matches_the_references_image(ctx, engine, tmp_path):
    image, variance, visible, n_walks = run_demo(engine, tmp_path, 1)
    ref = engine["ref_wost"][::-1]        # the fixture keeps the file's row order; Image::write stores the bottom row first
    # the same pixels are solved: the reference's image is zero exactly where the point is outside the domain or within the
    # boundary mask (a solved pixel of this problem is never exactly 0: the Dirichlet data are positive where walks end)
    ref_visible = ref != 0.0
    assert (visible != ref_visible).mean() < 2e-3, (visible != ref_visible).sum()
    both = visible & ref_visible
    assert both.mean() > 0.4
    # two independent 128-walk estimates of the same quantity: difference within 3 standard errors of the difference
    se = np.sqrt(2.0*variance[both]/n_walks)
    diff = (image - ref)[both]
    z = diff/np.maximum(se, 1e-4)
    assert np.mean(np.abs(diff) < 3.0*se + 1e-3) > 0.985, np.mean(np.abs(diff) < 3.0*se + 1e-3)
    noisy = se > 1e-3
    assert abs(float(np.mean(z[noisy]))) < 0.1 and 0.75 < float(np.std(z[noisy])) < 1.3, (np.mean(z[noisy]), np.std(z[noisy]))
    assert abs(float(np.mean(diff))) < 2e-3 and float(np.sqrt(np.mean(diff**2))) < 0.06
    pfm.write_pfm(str(tmp_path / "wost.pfm"), image)      # Image::write; reads back as the bottom-up file order
    assert np.array_equal(pfm.read_pfm(str(tmp_path / "wost.pfm"))[::-1, :, 0], image)


@pytest.mark.gpu
def test_demo_engine_four_channels_are_the_colour_solution(ctx, engine, tmp_path):
    """channels = 4 (runDemo<Array4>, demo.cpp:636-641): R, G, B of the colour textures and a zero channel on ONE set of walks; the
    luminance of that solution is the scalar demo's solution, walk for walk (same generators), and the zero channel stays zero"""
    colour, _, visible4, _ = run_demo(engine, tmp_path, 4)
    gray, _, visible1, _ = run_demo(engine, tmp_path, 1)
    assert np.array_equal(visible4, visible1)
    lum = 0.299*colour[..., 0] + 0.587*colour[..., 1] + 0.114*colour[..., 2]
    assert np.abs(lum - gray)[visible1].max() < 5e-5
    assert np.all(colour[..., 3] == 0.0)
    assert np.abs(colour[..., 0] - colour[..., 2])[visible1].max() > 0.01      # the channels do differ (the Dirichlet texture is nearly gray: 5/255 at most)


@pytest.mark.gpu
def test_demo_engine_boundary_value_caching_matches_the_references_image(ctx, engine, tmp_path):
    """engine/bvc.json against solutions/bvc.pfm. Both images are smooth (every pixel reads the same cache), so they differ by a
    smooth error field rather than by pixel noise: the reference's own bvc and rws images differ by 0.018 RMS, its bvc and wost
    images by 0.037 (the 128-walk noise of the latter)."""
    image, visible, solver = run_demo_bvc(engine, tmp_path)
    ref = engine["ref_bvc"][::-1]
    assert (visible != (ref != 0.0)).mean() < 2e-3
    both = visible & (ref != 0.0)
    diff = (image - ref)[both]
    # (the same pipeline over the compiled reference as backend, on the CPU: 0.017 RMS, mean -0.014 — one cache's common-mode error)
    assert float(np.sqrt(np.mean(diff**2))) < 0.035 and abs(float(diff.mean())) < 0.03, (np.sqrt(np.mean(diff**2)), diff.mean())
    # and against the reference's Walk on Stars image of the same problem: within its pixel noise
    wost = engine["ref_wost"][::-1]
    d2 = (image - wost)[both]
    assert float(np.sqrt(np.mean(d2**2))) < 0.05
    c = solver.caches
    assert len(c["absorbing"]) == 3072 and len(c["reflecting"]) == 3072 and "domain" not in c


def run_demo_rws(engine, tmp_path):
    """runReverseWalkOnStars (demo.cpp:330-454) on the problem of engine/rws.json: samples of the absorbing boundary displaced
    inward by five epsilon walk backwards and splat onto the evaluation grid (zb_rws_solve)"""
    from zombie_b200 import abi, bvc
    P = setup_problem(engine, tmp_path, 1, "rws")
    cfg, pde, queries, bbox, grid, where = (P[k] for k in ("cfg", "pde", "queries", "bbox", "grid", "where"))
    sc, mp = cfg["solver"], cfg["modelProblem"]
    eps_a = sc["epsilonShellForAbsorbingBoundary"]
    settings = zombie.Solvers.WalkSettings(eps_a, sc["epsilonShellForReflectingBoundary"], 1e-3, sc["russianRouletteThreshold"],
                                           sc["splittingThreshold"], sc["maxWalkLength"], 0, sc["maxWalkLength"], mp["solveDoubleSided"], False, False,
                                           False, sc["ignoreAbsorbingBoundaryContribution"], sc["ignoreReflectingBoundaryContribution"],
                                           sc["ignoreSourceContribution"], False)
    offset_a = 5.0*eps_a
    ap, ai, rp, ri = P["meshes"]
    inside_box = lambda p: np.all((p >= bbox[0]) & (p <= bbox[1]), axis=1)   # see run_demo_bvc
    sa = bvc.UniformBoundarySampler(ap, ai, inside_solve_region=inside_box, seed=31)
    sa.initialize(offset_a, mp["solveDoubleSided"])
    assert sc["ignoreReflectingBoundaryContribution"] and sc["ignoreSourceContribution"]   # -> those sample counts are 0 (:398, 416)
    s = sa.generate_samples(sc["absorbingBoundarySampleCount"], abi.ON_ABSORBING_BOUNDARY, offset_a)
    scene = queries.scene()
    da = scene.closest_point(abi.ABSORBING, s.pt, want_normal=False)[1]
    dr = scene.closest_point(abi.REFLECTING, s.pt, want_normal=False)[1]
    smp = abi.HostRwsSamples(s.pt, s.normal, s.type, s.pdf, da, dr, seed=0xE61E)
    ev = abi.HostRwsEvalPoints(grid[where])
    steps, records = scene.rws_solve(pde.descriptor(), settings, smp, ev, offset_a, radius_clamp=sc["radiusClampForKernels"], kernel_regularization=0.0)
    assert steps > 0 and records >= steps
    image, visible = to_image(P, ev.estimated_solution(smp.group_counts()))
    return image, visible


@pytest.mark.gpu
def test_demo_engine_reverse_walk_on_stars_matches_the_references_image(ctx, engine, tmp_path):
    """engine/rws.json (614,400 reverse walks) against solutions/rws.pfm; see the BVC test for the scale of the differences"""
    image, visible = run_demo_rws(engine, tmp_path)
    ref = engine["ref_rws"][::-1]
    assert (visible != (ref != 0.0)).mean() < 2e-3
    both = visible & (ref != 0.0)
    diff = (image - ref)[both]
    assert float(np.sqrt(np.mean(diff**2))) < 0.03 and abs(float(diff.mean())) < 0.015, (np.sqrt(np.mean(diff**2)), diff.mean())
    d2 = (image - engine["ref_wost"][::-1])[both]
    assert float(np.sqrt(np.mean(d2**2))) < 0.05
