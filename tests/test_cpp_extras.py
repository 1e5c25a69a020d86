"""bvc::BoundaryValueCaching::splat and rws::solve of the C++ mirror (include/zombie_b200/zombie.hpp) against the C ABI
they marshal to: the driver (tests/cpp/api_driver.cpp, 'extras' mode) turns the first 32 solved points into source
samples whose contribution is the walk estimate, splats them onto the first evaluation point and runs reverse walks
from them. The splat must equal zb_bvc_splat on the same cache (to float rounding of the bucket means); the reverse walks must be finite and
repeatable. Reference surfaces: boundary_value_caching.h:482-502, reverse_walk_solver.h:287-330."""
import subprocess

import numpy as np
import pytest

from tests import cases
from tests.test_cpp_api import CASES, driver, read_result, write_problem  # noqa: F401 (driver is a fixture)
from zombie_b200 import abi, native


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["wost3d_robin_grad", "wost2d_neumann"])
def test_cpp_bvc_and_rws_surface(ctx, driver, tmp_path, name):  # noqa: F811
    c = CASES[name]
    dim = c["dim"]
    whole, absorbing, reflecting = cases.mesh3d(3) if dim == 3 else cases.mesh2d(512)
    if c["kind"] == abi.REFLECTING_NONE:
        absorbing, reflecting = whole, None
    pts = cases.query_points(256, dim, 0.6, seed=17)
    prob, out = str(tmp_path / "p.bin"), str(tmp_path / "o.bin")
    write_problem(prob, dim, c["estimator"], c["kind"], absorbing, reflecting, pts, 64, c["quantity"], c["family"],
                  c["params"], c["sigma"], c["mu"])
    runs = []
    for _ in range(2):
        r = subprocess.run([driver, prob, out, "extras"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        assert r.returncode == 0, r.stderr
        runs.append(read_result(out, len(pts)))
    steps, rec, (bvc_cpp, rws_cpp) = runs[0]
    assert np.isfinite([bvc_cpp, rws_cpp]).all()
    assert runs[1][2] == (bvc_cpp, rws_cpp)          # same seeds, same record order, no atomics: repeatable

    # the same splat straight through the C ABI
    if reflecting is None:
        sc = native.Scene(ctx, dim, absorbing=absorbing)
    else:
        sc = native.Scene(ctx, dim, **cases.scene_kwargs(dim, c["kind"], absorbing, reflecting, mu=c["mu"]))
    _, da, _, _ = sc.closest_point(abi.ABSORBING, pts)
    dr = sc.closest_point(abi.REFLECTING, pts)[1] if reflecting is not None else np.full(len(pts), np.inf, np.float32)
    k = 32
    z = np.zeros(k, np.float32)
    samples = abi.HostBvcSamples(pts[:k], np.zeros((k, dim), np.float32), np.full(k, abi.IN_DOMAIN, np.int32),
                                 np.ones(k, np.float32), z, z, rec[:k, 0], z)
    ev = abi.HostBvcEvalPoints(pts[:1], da[:1], dr[:1])
    ctx.bvc_splat(dim, c["sigma"], samples, ev, radius_clamp=1e-3)
    # the mirror re-derives the estimate from the returned (count, sum) buckets: same numbers, host rounding of the mean
    assert abs(bvc_cpp - ev.solution[0]) <= 2e-6*max(1.0, abs(ev.solution[0]))
