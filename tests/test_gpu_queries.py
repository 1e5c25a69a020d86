"""GPU parity tests of the geometric queries, through the C ABI (zb_closest_point, zb_intersect_ray,
zb_occluded, zb_star_radius, zb_sample_sphere, zb_pcg32_*), against (a) the committed golden fixtures that
the reference produced (tests/golden/make_golden.py) and (b) the compiled reference itself on larger seeded
inputs. Gates (BASELINE.json north_star): primitive index identical, distance and point within 1e-6 relative."""
import json
import os

import numpy as np
import pytest

from tests import cases
from tests.conftest import GOLDEN, rel_err
from zombie_b200 import abi, native, procedural as P

pytestmark = pytest.mark.gpu
TOL = 1e-6


def check_closest(ref_out, gpu_out, x, allow_tie_fraction=0.5):
    rprim, rd, rpt, rn = ref_out
    gprim, gd, gpt, gn = gpu_out
    assert np.array_equal(rprim >= 0, gprim >= 0)
    found = rprim >= 0
    assert rel_err(rd, gd)[found].max() <= TOL
    mism = (rprim != gprim) & found
    # Against the reference's BVH a different index is only acceptable for an EXACT tie: the query is bit-exactly
    # equidistant from two primitives — almost always two primitives sharing the closest edge/vertex (same closest
    # point), rarely two unrelated ones whose float distances coincide. FCPW keeps whichever primitive it visits last
    # (d*d <= r2 overwrites): its BVH answers by tree topology (bvh.inl:1331-1335), its brute-force Baseline by input
    # order, i.e. with the LARGEST tied index. zb_closest_point reports the Baseline's answer on every query
    # (test_gpu_queries_vs_reference.py::test_closest_point_index_is_the_reference_baselines asserts that with no
    # allowance); here, against BVH-generated goldens, a mismatch must therefore be an exact tie won by the larger index.
    assert np.array_equal(rd[mism], gd[mism]), "primitive index differs without an exact distance tie"
    assert (gprim[mism] > rprim[mism]).all(), "a tie must go to the largest input index"
    assert mism.mean() <= allow_tie_fraction
    ok = found & ~mism
    scale = np.maximum(1.0, np.abs(rpt).max())
    assert np.abs(rpt - gpt)[ok].max() <= TOL*scale
    assert np.abs(np.linalg.norm(x - gpt, axis=1) - gd)[mism].max(initial=0.0) <= 2e-6*scale
    if rn is not None and gn is not None:
        assert np.abs(rn - gn)[ok].max() <= 1e-5


def check_rays(ref_out, gpu_out):
    rprim, rt, rpt, rn = ref_out
    gprim, gt, gpt, gn = gpu_out
    assert np.array_equal(rprim >= 0, gprim >= 0)
    hit = rprim >= 0
    assert rel_err(rt, gt)[hit].max() <= TOL
    mism = (rprim != gprim) & hit
    assert np.array_equal(rt[mism], gt[mism])
    assert mism.mean() <= 0.01
    ok = hit & ~mism
    assert np.abs(rpt - gpt)[ok].max() <= TOL*np.maximum(1.0, np.abs(rpt).max())
    assert np.abs(rn - gn)[ok].max() <= 1e-5


def test_pcg32_known_answer_vector_on_device(ctx):
    g = json.load(open(os.path.join(GOLDEN, "pcg32_demo.json")))
    u, coins, rolls, cards = ctx.pcg32_demo(g["initstate"], g["initseq"], len(g["rounds"]))
    for r, rd in enumerate(g["rounds"]):
        assert [int(v) for v in u[r]] == rd["u32"]
        assert [int(v) for v in coins[r]] == rd["coins"]
        assert [int(v) for v in rolls[r]] == rd["rolls"]
        assert [int(v) for v in cards[r]] == rd["cards"]
    assert [int(v) for v in ctx.pcg32(42, 54, 6)] == g["rounds"][0]["u32"]


def test_pcg32_float_stream_matches_reference(ctx, ref):
    assert np.array_equal(ctx.pcg32_floats(0xC0FFEE, 12345, 4096), ref.pcg32_floats(0xC0FFEE, 12345, 4096))
    assert np.array_equal(ctx.pcg32(7, 2**40 + 3, 1000), ref.pcg32(7, 2**40 + 3, 1000))


@pytest.mark.parametrize("dim", [2, 3])
def test_queries_match_golden_fixtures(ctx, dim):
    g = np.load(os.path.join(GOLDEN, "queries_%dd.npz" % dim))
    whole, absorbing, reflecting = cases.mesh3d(3) if dim == 3 else cases.mesh2d(512)
    nq = len(g["cp_a_prim"])
    x = cases.query_points(nq, dim, 0.97, seed=3)
    o, d, tmax = cases.rays(nq, dim, 0.6, seed=5)
    xs = cases.query_points(nq, dim, 0.99, seed=9)
    minr = np.full(nq, 1e-3, np.float32)
    for kind, name in ((abi.REFLECTING_NEUMANN, "neumann"), (abi.REFLECTING_ROBIN, "robin")):
        sc = native.Scene(ctx, dim, **cases.scene_kwargs(dim, kind, absorbing, reflecting, mu=1.0))
        if kind == abi.REFLECTING_NEUMANN:
            for which, w in ((abi.ABSORBING, "a"), (abi.REFLECTING, "r")):
                check_closest((g["cp_%s_prim" % w], g["cp_%s_dist" % w], g["cp_%s_pt" % w], g["cp_%s_normal" % w]),
                              sc.closest_point(which, x), x)
                prim, dist, _, _ = sc.closest_point(which, x, np.full(nq, 0.15, np.float32))
                assert np.array_equal(prim >= 0, g["cpr_%s_prim" % w] >= 0)
                assert (prim < 0).any() and (prim >= 0).any()
                f = prim >= 0
                assert rel_err(dist, g["cpr_%s_dist" % w])[f].max() <= TOL
                check_rays((g["ray_%s_prim" % w], g["ray_%s_dist" % w], g["ray_%s_pt" % w], g["ray_%s_normal" % w]),
                           sc.intersect_ray(which, o, d, tmax))
                assert np.array_equal(sc.occluded(which, o, o + d*tmax[:, None]), g["occ_%s" % w])
        da = g["star_maxr"]
        for key, flip in (("star_%s" % name, None), ("star_%s_flip" % name, np.ones(nq, np.uint8))):
            r = sc.star_radius(xs, minr, da, flip=flip)
            assert np.abs(r - g[key]).max() <= 1e-3*max(1.0, float(g[key].max())), key   # FCPW's silhouette tolerance
            assert np.mean(rel_err(r, g[key]) <= 1e-5) > 0.99, key
        sc.close()


@pytest.mark.parametrize("dim", [2, 3])
def test_sphere_sampling_is_a_valid_importance_sampler(ctx, ref, dim):
    """The stochastic descent's pdf depends on the tree topology (SURVEY.md hard part 4), so parity is
    statistical: E[1/pdf] = area of the boundary inside the sphere for both trees; samples lie inside the
    sphere, on the returned primitive, and the reported pdf is positive."""
    whole, absorbing, reflecting = cases.mesh3d(4) if dim == 3 else cases.mesh2d(2048)
    kw = cases.scene_kwargs(dim, abi.REFLECTING_NEUMANN, absorbing, reflecting)
    rs, gs = ref.RefScene(dim, **kw), native.Scene(ctx, dim, **kw)
    n = 100000
    x = np.repeat(cases.query_points(50, dim, 0.8, seed=51), n//50, axis=0)
    rad = np.full(n, 0.6, np.float32)
    u = np.random.RandomState(5).uniform(size=(n, dim)).astype(np.float32)
    rprim, rpt, rn, rpdf = rs.sample_sphere(x, rad, u)
    gprim, gpt, gn, gpdf = gs.sample_sphere(x, rad, u)
    # (how often the single-path descent dead-ends depends on the tree; the pdf accounts for it)
    f = gprim >= 0
    assert (gpdf[f] > 0).all()
    # a sample is a uniform point of a primitive that touches the sphere, so it may lie a primitive's diameter
    # outside it (the estimator rejects those, walk_on_stars.h:222); both trees put the same mass inside
    rf = rprim >= 0
    edge = np.linalg.norm(reflecting[0][reflecting[1][:, 0]] - reflecting[0][reflecting[1][:, 1]], axis=1).max()
    assert (np.linalg.norm(gpt - x, axis=1)[f] <= 0.6 + 1.01*edge).all()
    inside_g = (np.linalg.norm(gpt - x, axis=1)[f] <= 0.6).mean()
    inside_r = (np.linalg.norm(rpt - x, axis=1)[rf] <= 0.6).mean()
    assert abs(inside_g - inside_r) < 0.05, (inside_g, inside_r)
    # per query point: area estimates agree within 4 standard errors
    bad = 0
    for j in range(50):
        s = slice(j*(n//50), (j + 1)*(n//50))
        wr = np.where(rprim[s] >= 0, 1.0/np.maximum(rpdf[s], 1e-30), 0.0)
        wg = np.where(gprim[s] >= 0, 1.0/np.maximum(gpdf[s], 1e-30), 0.0)
        se = np.sqrt(wr.var()/len(wr) + wg.var()/len(wg))
        bad += abs(wr.mean() - wg.mean()) > 4.0*se + 1e-6
    assert bad <= 2
    # the sampled point lies on the primitive it names, with that primitive's normal
    pos, idx = reflecting
    a = pos[idx[gprim[f], 0]]
    nrm = gn[f]
    assert np.abs(((gpt[f] - a)*nrm).sum(axis=1)).max() < 1e-5
    rs.close(); gs.close()


def test_edge_cases(ctx):
    whole, absorbing, reflecting = cases.mesh3d(2)
    sc = native.Scene(ctx, 3, **cases.scene_kwargs(3, abi.REFLECTING_NEUMANN, absorbing, reflecting))
    # empty batches are fine
    e3 = np.zeros((0, 3), np.float32)
    prim, dist, pt, nrm = sc.closest_point(abi.ABSORBING, e3)
    assert len(prim) == 0
    assert len(sc.occluded(abi.REFLECTING, e3, e3)) == 0
    # ragged batch sizes (not a multiple of the warp/CTA size)
    for n in (1, 31, 33, 129):
        x = cases.query_points(n, 3, 0.9, seed=n)
        prim, dist, pt, nrm = sc.closest_point(abi.ABSORBING, x)
        assert (prim >= 0).all() and np.allclose(np.linalg.norm(x - pt, axis=1), dist, rtol=1e-5, atol=1e-7)
    # nothing within the radius -> -1 / FLT_MAX like fcpw::Interaction's defaults
    x = np.zeros((4, 3), np.float32)
    prim, dist, _, _ = sc.closest_point(abi.ABSORBING, x, np.full(4, 1e-3, np.float32))
    assert (prim == -1).all() and (dist == np.float32(abi.FLT_MAX)).all()
    # rays that miss
    o = np.full((3, 3), 5.0, np.float32)
    d = np.tile(np.array([[1.0, 0, 0]], np.float32), (3, 1))
    prim, t, _, _ = sc.intersect_ray(abi.REFLECTING, o, d, None)
    assert (prim == -1).all()
    # a scene without a reflecting boundary refuses reflecting queries with an error code, not a crash
    sc2 = native.Scene(ctx, 3, absorbing=whole)
    with pytest.raises(native.ZombieError, match="empty"):
        sc2.closest_point(abi.REFLECTING, x)
    with pytest.raises(native.ZombieError):
        native.Scene(ctx, 4, absorbing=whole)
    sc.close(); sc2.close()


def test_full_size_mesh_properties(ctx):
    """BASELINE-size boundary (1,310,720 triangles): size-independent properties of the closest-point query —
    the returned point is at the returned distance, lies on the named triangle, no triangle of a random
    subsample is closer, and the result is idempotent (the closest point of a closest point is itself)."""
    pos, tri = P.bumpy_sphere(8)
    sc = native.Scene(ctx, 3, absorbing=(pos, tri))
    st = sc.stats(abi.ABSORBING)
    assert st["nodes"] > 40000 and st["bytes"] < 200e6
    x = cases.query_points(1 << 20, 3, 0.98, seed=61)
    prim, dist, pt, nrm = sc.closest_point(abi.ABSORBING, x)
    assert (prim >= 0).all() and prim.max() < len(tri)
    assert np.abs(np.linalg.norm(x.astype(np.float64) - pt, axis=1) - dist).max() < 2e-6
    a, b, c = (pos[tri[prim, k]].astype(np.float64) for k in range(3))
    n = np.cross(b - a, c - a)
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    assert np.abs(((pt - a)*n).sum(axis=1)).max() < 1e-5
    # vertices of any triangle are never closer than the reported distance
    sub = np.random.RandomState(1).choice(len(pos), 4096, replace=False)
    for j in range(0, 64):
        dv = np.linalg.norm(pos[sub].astype(np.float64) - x[j], axis=1).min()
        assert dv >= dist[j] - 1e-6
    prim2, dist2, pt2, _ = sc.closest_point(abi.ABSORBING, pt[:4096])
    assert dist2.max() < 1e-6
    sc.close()
