"""Small end-to-end case for compute-sanitizer (memcheck / racecheck): device-built hierarchies in 2D and 3D (large enough for the
CTA-per-slot cone pass), a vertex update, batched queries, a scalar solve and a three-channel solve.
    compute-sanitizer --tool memcheck python scripts/sanitize_case.py"""
import ctypes as C
import sys

import numpy as np

sys.path.insert(0, ".")
from tests import cases  # noqa: E402
from zombie_b200 import abi, native, workloads  # noqa: E402

ctx = native.Context(0)
for dim, size in ((3, 5), (2, 8192)):
    whole, absorbing, reflecting = cases.mesh3d(size) if dim == 3 else cases.mesh2d(size)
    kw = cases.scene_kwargs(dim, abi.REFLECTING_ROBIN, absorbing, reflecting, mu=1.0)
    s = native.Scene(ctx, dim, builder=abi.BVH_BUILDER_DEVICE, **kw)
    assert s.validate(abi.ABSORBING, *absorbing) == 0 and s.validate(abi.REFLECTING, *reflecting) == 0
    s.update_vertices(abi.REFLECTING, (reflecting[0]*1.01).astype(np.float32))
    s.update_robin_coefficients(np.full(len(reflecting[1]), 0.5, np.float32), np.full(len(reflecting[1]), 0.7, np.float32))
    x = cases.query_points(2000, dim, 0.9, seed=1)
    _, da, _, _ = s.closest_point(abi.ABSORBING, x)
    s.star_radius(x, np.full(len(x), 1e-3, np.float32), da)
    s.close()
w = workloads.config3(subdivisions=3, n_points=128, n_walks=8)
s = native.Scene(ctx, 3, builder=abi.BVH_BUILDER_DEVICE, **w.scene_kwargs())
_, da, _, _ = s.closest_point(abi.ABSORBING, w.points)
_, dr, _, _ = s.closest_point(abi.REFLECTING, w.points)
pdes = []
for k in range(3):
    d = abi.PdeDesc()
    C.memmove(C.byref(d), C.byref(w.pde), C.sizeof(abi.PdeDesc))
    d.params[3] = 1.0 + k
    pdes.append(d)
pts = abi.HostSamplePoints(w.points, da, dr, nWalks=w.n_walks, estimationQuantity=w.quantity)
stats = [abi.HostStatistics(pts.n, 3) for _ in pdes]
steps = s.solve_channels(w.estimator, pdes, w.settings, pts, stats)
assert steps > 0 and np.isfinite(stats[2].solutionMean).all()
assert np.allclose(stats[1].solutionMean, 2.0*stats[0].solutionMean, rtol=1e-4, atol=1e-5)   # the family is linear in params[3]
print("sanitize case ok:", steps, "walk steps")
