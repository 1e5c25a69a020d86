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
        # The ONLY accepted cause: FCPW's SNCH traversal occasionally prunes a node that holds an edge its own primitive
        # test classifies as a silhouette through the `precision` rules (silhouette_edges.inl:88-105), so the reference's
        # BVH deviates from the reference's brute-force Baseline on a few queries in 10^5 (by more than FCPW's own 1e-3
        # tolerance). The device traversal treats nodes within `precision` of the query as candidates and must agree
        # with that Baseline on every such row (no rate cap: each row is checked). The Robin radius bound itself is
        # evaluated in the reference's operation order without FMA contraction (queries.cuh: robinRadiusBound), so
        # queries nearly coplanar with a primitive, where h^2 mu/(c(h mu - c^2)) cancels, agree as well.
        rb = ref.RefScene(dim, use_baseline=True, **kw)
        base = rb.star_radius(x[bad], minr[bad], da[bad])
        assert (np.abs(base - gr[bad]) <= 1e-5*max(1.0, float(rr.max()))).all(), (x[bad], rr[bad], gr[bad], base)
        rb.close()
    assert np.mean(rel_err(rr, gr)[~bad] <= 1e-4) > 0.99   # the Robin radius bound h^2 mu/(c(h mu - c^2)) amplifies float rounding
    assert (rr < 0.999*da).mean() > 0.01   # the query is exercised: some radii are limited by silhouettes / reflectance
    rs.close(); gs.close()
