"""GPU parity of every geometric query against the COMPILED reference on large seeded inputs (200,000 3D / 50,000 2D queries
per configuration): closest point, closest ray hit, occlusion, star radius. The golden-fixture query tests are in
test_gpu_queries.py."""
import numpy as np
import pytest

from tests import cases
from tests.conftest import rel_err
from tests.test_gpu_queries import check_closest, check_rays
from zombie_b200 import abi, native

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dim,kind", [(3, abi.REFLECTING_ROBIN), (3, abi.REFLECTING_NEUMANN), (2, abi.REFLECTING_NEUMANN),
                                      (2, abi.REFLECTING_ROBIN)])
def test_queries_match_compiled_reference(ctx, ref, dim, kind):
    nq = 200000 if dim == 3 else 50000
    whole, absorbing, reflecting = cases.mesh3d(5) if dim == 3 else cases.mesh2d(4096)
    kw = cases.scene_kwargs(dim, kind, absorbing, reflecting, mu=2.0)
    rs, gs = ref.RefScene(dim, **kw), native.Scene(ctx, dim, **kw)
    x = cases.query_points(nq, dim, 0.99, seed=41)
    o, d, tmax = cases.rays(nq, dim, 0.7, seed=43)
    for which in (abi.ABSORBING, abi.REFLECTING):
        check_closest(rs.closest_point(which, x), gs.closest_point(which, x), x)
        check_rays(rs.intersect_ray(which, o, d, tmax), gs.intersect_ray(which, o, d, tmax))
        b = o + d*tmax[:, None]
        assert np.array_equal(rs.occluded(which, o, b), gs.occluded(which, o, b))
    _, da, _, _ = rs.closest_point(abi.ABSORBING, x)
    minr = np.full(nq, 1e-3, np.float32)
    rr, gr = rs.star_radius(x, minr, da), gs.star_radius(x, minr, da)
    tol = 1e-3*max(1.0, float(rr.max()))
    bad = np.abs(rr - gr) > tol
    if bad.any():
        # Two causes, each asserted row by row (no rate cap):
        # (1) FCPW's SNCH traversal occasionally prunes a node that holds an edge its own primitive test classifies as a
        #     silhouette through the `precision` rules (silhouette_edges.inl:88-105), so the reference's BVH deviates from
        #     the reference's brute-force Baseline on a few queries in 10^5 (by more than FCPW's own 1e-3 tolerance). The
        #     device traversal treats nodes within `precision` of the query as candidates and must agree with that Baseline.
        # (2) Robin only: the query point lies IN the plane of a reflecting primitive whose distance is about the radius
        #     (plane distance h below 1e-5 of the distance d). The reference's radius for that primitive is
        #     h^2 mu/(c (h mu - c^2)) with h and the cosine c = h/d both at the rounding level of the dot products they come
        #     from (robin_bounds.h:48-74, reflectance_boundary_bvh/geometry.h:196-226) and comes out up to a few percent
        #     BELOW d, which no exact evaluation could. Whether that value counts then depends on the visiting order: the
        #     primitive is skipped once the radius has fallen below its squared distance (geometry.h:179-182), so the
        #     reference's BVH, its Baseline and any other traversal may each report a different radius — they do: on the
        #     1.3 M-triangle mesh the reference's BVH and Baseline disagree with each other on such rows by the same 0.5-1 %
        #     (the device keeps the reference's arithmetic and early exit, so it is one more such traversal; one query per
        #     launch reproduces the reference's or the Baseline's value on every such row, tests/diagnostics/star_fullsize_debug.py).
        rb = ref.RefScene(dim, use_baseline=True, **kw)
        base = rb.star_radius(x[bad], minr[bad], da[bad])
        rb.close()
        by_baseline = np.abs(base - gr[bad]) <= 1e-5*max(1.0, float(rr.max()))
        coplanar = np.zeros(int(bad.sum()), bool)
        if kind == abi.REFLECTING_ROBIN:
            rp, ri = reflecting
            for k, i in enumerate(np.where(bad)[0]):
                coplanar[k] = cases.coplanar_primitive_within(rp.astype(np.float64), ri, x[i].astype(np.float64),
                                                              0.97*min(rr[i], gr[i]), 1.15*max(rr[i], gr[i]), 1e-5)
        assert (by_baseline | coplanar).all(), (x[bad], rr[bad], gr[bad], base, by_baseline, coplanar)
        assert np.abs(rr - gr)[bad][coplanar & ~by_baseline].max(initial=0.0) <= 0.1*float(rr.max())
        assert coplanar.sum() <= 2 + 5e-4*nq   # a thin set of configurations: a handful of rows
    assert np.mean(rel_err(rr, gr)[~bad] <= 1e-4) > 0.99   # the Robin radius bound h^2 mu/(c(h mu - c^2)) amplifies float rounding
    assert (rr < 0.999*da).mean() > 0.01   # the query is exercised: some radii are limited by silhouettes / reflectance
    rs.close(); gs.close()


@pytest.mark.parametrize("dim", [3, 2])
def test_closest_point_index_is_the_reference_baselines(ctx, ref, dim):
    """Primitive index, bit for bit. When a query is exactly equidistant from several primitives (closest point on a shared
    edge or vertex) FCPW keeps the one it visits LAST (d*d <= r2 replaces, triangles.inl:479): the reference's BVH answers
    by its tree topology, the reference's brute-force Baseline aggregate (baseline.inl:349-372) by input order — the largest
    index among the tied primitives. zb_closest_point states that rule (include/zombie_b200.h) and must therefore equal the
    Baseline on EVERY query: index, distance, point and normal, with no tie allowance. The reference's BVH must then agree
    except on exact ties, where its distance and point are still identical."""
    whole, absorbing, reflecting = cases.mesh3d(4) if dim == 3 else cases.mesh2d(2048)
    kw = cases.scene_kwargs(dim, abi.REFLECTING_NEUMANN, absorbing, reflecting)
    nq = 20000
    x = cases.query_points(nq, dim, 0.99, seed=241)
    # queries ON vertices and edge midpoints of the boundary and along vertex normals: ties of 2 up to 6 primitives
    pos, idx = absorbing
    rng = np.random.RandomState(7)
    v = pos[idx[rng.randint(0, len(idx), 2000), 0]]
    e = 0.5*(pos[idx[:2000, 0]] + pos[idx[:2000, 1]])
    x = np.concatenate([x, v, (v*np.float32(0.75)).astype(np.float32), e]).astype(np.float32)
    rb, rs, gs = ref.RefScene(dim, use_baseline=True, **kw), ref.RefScene(dim, **kw), native.Scene(ctx, dim, **kw)
    ties = 0
    for which in (abi.ABSORBING, abi.REFLECTING):
        bprim, bd, bpt, bn = rb.closest_point(which, x)
        gprim, gd, gpt, gn = gs.closest_point(which, x)
        rprim, rd, rpt, rn = rs.closest_point(which, x)
        assert np.array_equal(bprim, gprim)
        assert np.array_equal(bd, gd) and np.array_equal(bpt, gpt)
        assert np.abs(bn - gn).max() <= 1e-6
        other = rprim != gprim
        ties += int(other.sum())
        assert np.array_equal(rd[other], gd[other]) and (rprim[other] < gprim[other]).all()
        # bounded searches follow the same rule
        rad = np.full(len(x), 0.2, np.float32)
        bprim, bd, _, _ = rb.closest_point(which, x, rad)
        gprim, gd, _, _ = gs.closest_point(which, x, rad)
        assert np.array_equal(bprim, gprim) and np.array_equal(bd, gd) and (bprim < 0).any() and (bprim >= 0).any()
    assert ties > 100   # the rule is exercised
    rb.close(); rs.close(); gs.close()


@pytest.mark.parametrize("dim", [3, 2])
def test_all_hits_counts_and_inside_domain_match_compiled_reference(ctx, ref, dim):
    """zb_count_ray_hits = Aggregate::intersect(ray, interactions, false, true) as a count (bvh.inl:890-1028: every hit,
    sorted, duplicates on shared edges merged) and zb_inside_domain = GeometricQueries::insideDomain
    (geometric_queries.h:357-371: the parity of that count along each axis over both boundaries), on a closed domain split
    into an absorbing and a reflecting part; query points inside, outside and (2D) in the concave bays of the star."""
    whole, absorbing, reflecting = cases.mesh3d(4) if dim == 3 else cases.mesh2d(2048)
    kw = cases.scene_kwargs(dim, abi.REFLECTING_NEUMANN, absorbing, reflecting)
    rs, gs = ref.RefScene(dim, **kw), native.Scene(ctx, dim, **kw)
    nq = 50000
    x = (cases.query_points(nq, dim, 1.0, seed=77)*np.float32(1.6)).astype(np.float32)
    # points whose axis rays pass exactly through mesh vertices / along edges: the duplicate-merging rule decides
    pos = whole[0]
    v = pos[np.random.RandomState(3).randint(0, len(pos), 2000)].copy()
    v[:, 0] -= np.float32(0.37)
    x = np.concatenate([x, v]).astype(np.float32)
    o, d, tmax = cases.rays(nq, dim, 1.2, seed=79)
    for which in (abi.ABSORBING, abi.REFLECTING):
        rc, gc = rs.count_ray_hits(which, o, d, tmax*3), gs.count_ray_hits(which, o, d, tmax*3)
        assert np.array_equal(rc, gc)
        assert (rc >= 2).any() and (rc == 0).any()
        ex = np.zeros_like(x)
        ex[:, 0] = 1.0
        assert np.array_equal(rs.count_ray_hits(which, x, ex), gs.count_ray_hits(which, x, ex))
    ri, gi = rs.inside_domain(x), gs.inside_domain(x)
    assert np.array_equal(ri, gi)
    assert 0.1 < ri.mean() < 0.9
    # a domain declared not watertight is "inside" everywhere (geometric_queries.h:358)
    open_scene = native.Scene(ctx, dim, watertight=False, **kw)
    assert open_scene.inside_domain(x[:100]).all()
    open_scene.close(); rs.close(); gs.close()


@pytest.mark.parametrize("dim", [3, 2])
def test_robin_coefficient_refit_on_the_device(ctx, ref, dim):
    """zb_scene_update_robin_coefficients(_dev) = FcpwRobinBoundaryHandler::updateCoefficientValues
    (fcpw_geometric_queries.h:997-1010): after the device-side refit of the per-primitive records and of every node's
    coefficient range, star radii equal those of a scene BUILT with the new coefficients bit for bit, and those of the
    reference after its own update."""
    import torch
    whole, absorbing, reflecting = cases.mesh3d(4) if dim == 3 else cases.mesh2d(2048)
    n_r = len(reflecting[1])
    rng = np.random.RandomState(17)
    mu_old = np.full(n_r, 0.3, np.float32)
    mn_new = rng.uniform(0.5, 4.0, n_r).astype(np.float32)
    mx_new = (mn_new*rng.uniform(1.0, 1.5, n_r)).astype(np.float32)
    kw_old = cases.scene_kwargs(dim, abi.REFLECTING_ROBIN, absorbing, reflecting, mu=1.0)
    kw_old["min_robin"], kw_old["max_robin"] = mu_old, mu_old
    kw_new = dict(kw_old, min_robin=mn_new, max_robin=mx_new)
    nq = 20000
    x = cases.query_points(nq, dim, 0.99, seed=91)
    minr = np.full(nq, 1e-3, np.float32)
    updated, fresh = native.Scene(ctx, dim, **kw_old), native.Scene(ctx, dim, **kw_new)
    _, da, _, _ = fresh.closest_point(abi.ABSORBING, x)
    before = updated.star_radius(x, minr, da)
    want = fresh.star_radius(x, minr, da)
    assert (before != want).mean() > 0.05                      # the coefficients matter on this problem
    updated.update_robin_coefficients(mn_new, mx_new)          # host arrays
    assert np.array_equal(updated.star_radius(x, minr, da), want)
    updated.update_robin_coefficients(mu_old, mu_old)
    assert np.array_equal(updated.star_radius(x, minr, da), before)
    dev = torch.device("cuda", ctx.device)
    tmn, tmx = torch.from_numpy(mn_new).to(dev), torch.from_numpy(mx_new).to(dev)
    torch.cuda.synchronize(dev)
    updated.update_robin_coefficients_dev(tmn.data_ptr(), tmx.data_ptr())   # device arrays, asynchronous
    assert np.array_equal(updated.star_radius(x, minr, da), want)
    # and a solve sees the new coefficients: same estimates as on the freshly built scene (deterministic scheduler)
    pde, _ = cases.exp_pde(mu=1.0, sigma=1.0)
    xs = x[:512]
    out = []
    for sc in (updated, fresh):
        _, dr, _, _ = sc.closest_point(abi.REFLECTING, xs)
        pts = abi.HostSamplePoints(xs, da[:512], dr, nWalks=16, estimationQuantity=abi.SOLUTION)
        st = abi.HostStatistics(len(xs), dim)
        sc.solve(abi.WALK_ON_STARS, pde, abi.WalkSettings.default(), pts, st)
        out.append(st.solutionMean.copy())
    assert np.array_equal(out[0], out[1])
    rs = ref.RefScene(dim, **kw_new)
    rr = rs.star_radius(x, minr, da)
    assert np.mean(np.abs(rr - want) <= 1e-3*max(1.0, float(rr.max()))) > 0.999
    updated.close(); fresh.close(); rs.close()


def test_full_size_mesh_matches_compiled_reference(ctx, ref):
    """The BASELINE headline geometry itself — the 1,310,720-triangle bumpy icosphere of config C3, Robin z > 0 /
    Dirichlet z < 0 — against the compiled reference: 20,000 queries of every kind (closest point on both boundaries, first
    hit, occlusion, Robin star radius) and a 256-point x 16-walk solve, with the device status word clean afterwards (no work
    list or traversal stack overflowed in the pooled phase, which is exercised hardest at this size).
    Mirrors deps/fcpw/tests/aggregate_tests.cpp:393,411,486,562,610 at the benchmark's size."""
    from zombie_b200 import workloads
    nq = 20000
    w = workloads.config3(subdivisions=8, n_points=256, n_walks=16)
    kw = w.scene_kwargs()
    assert len(kw["absorbing"][1]) + len(kw["reflecting"][1]) == 1310720
    rs, gs = ref.RefScene(3, **kw), native.Scene(ctx, 3, **kw)
    x = cases.query_points(nq, 3, 0.95, seed=141)
    o, d, tmax = cases.rays(nq, 3, 0.8, seed=143)
    for which in (abi.ABSORBING, abi.REFLECTING):
        check_closest(rs.closest_point(which, x), gs.closest_point(which, x), x)
        check_rays(rs.intersect_ray(which, o, d, tmax), gs.intersect_ray(which, o, d, tmax))
        b = o + d*tmax[:, None]
        assert np.array_equal(rs.occluded(which, o, b), gs.occluded(which, o, b))
    _, da, _, _ = rs.closest_point(abi.ABSORBING, x)
    minr = np.full(nq, 1e-3, np.float32)
    rr, gr = rs.star_radius(x, minr, da), gs.star_radius(x, minr, da)
    tol = 1e-3*max(1.0, float(rr.max()))
    bad = np.abs(rr - gr) > tol
    if bad.any():   # the two causes asserted row by row in test_queries_match_compiled_reference
        rb = ref.RefScene(3, use_baseline=True, **kw)
        base = rb.star_radius(x[bad], minr[bad], da[bad])
        rb.close()
        by_baseline = np.abs(base - gr[bad]) <= 1e-5*max(1.0, float(rr.max()))
        rp, ri = kw["reflecting"]
        coplanar = np.array([cases.coplanar_primitive_within(rp.astype(np.float64), ri, x[i].astype(np.float64),
                                                             0.97*min(rr[i], gr[i]), 1.15*max(rr[i], gr[i]), 1e-5)
                             for i in np.where(bad)[0]], bool)
        assert (by_baseline | coplanar).all(), (x[bad], rr[bad], gr[bad], base)
        assert np.abs(rr - gr)[bad].max() <= 0.1*float(rr.max())
        assert bad.sum() <= 2 + 5e-4*nq
    assert np.mean(rel_err(rr, gr)[~bad] <= 1e-4) > 0.99
    assert (rr < 0.999*da).mean() > 0.01

    # a short solve on the same scene: estimates within 3 standard errors of the reference's, same walk lengths
    _, pa, _, _ = gs.closest_point(abi.ABSORBING, w.points)
    _, pr, _, _ = gs.closest_point(abi.REFLECTING, w.points)
    res = []
    for solver, seed in ((gs, 0xBEEF), (rs, 0xF00D)):
        pts = abi.HostSamplePoints(w.points, pa, pr, nWalks=w.n_walks, estimationQuantity=abi.SOLUTION, seed=seed)
        st = abi.HostStatistics(pts.n, 3)
        out = solver.solve(w.estimator, w.pde, w.settings, pts, st)
        res.append((st, out if np.isscalar(out) else out[0]))
    z = cases.z_scores(res[0][0], res[1][0])
    assert np.mean(np.abs(z) < 3.0) >= 0.98 and abs(z.mean()) < 0.25, (np.mean(np.abs(z) < 3.0), z.mean())
    assert abs(res[0][1]/float(res[1][1]) - 1.0) < 0.1
    # every call above returned ZB_OK: the C ABI reads the device status word after each entry point and fails with
    # "traversal stack overflow" / "cooperative traversal stack overflow" when a kernel raised it (capi.cu: checkStatus)
    rs.close(); gs.close()
