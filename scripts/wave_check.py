"""Diagnostic: estimates per point and run-to-run repeatability of the wavefront scheduler on a small WoSt and WoS solve."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests import cases
from zombie_b200 import abi, native

ctx = native.Context(0)
for name, dim, est, kind, n in (("wost3d", 3, abi.WALK_ON_STARS, abi.REFLECTING_NEUMANN, 2048), ("wos3d", 3, abi.WALK_ON_SPHERES, abi.REFLECTING_NONE, 2048)):
    whole, absorbing, reflecting = cases.mesh3d(3)
    sc = native.Scene(ctx, 3, absorbing=whole) if kind == abi.REFLECTING_NONE else native.Scene(ctx, 3, **cases.scene_kwargs(3, kind, absorbing, reflecting, mu=1.0))
    x = cases.query_points(n, 3, 0.7, seed=13)
    _, da, _, _ = sc.closest_point(abi.ABSORBING, x)
    dr = None if kind == abi.REFLECTING_NONE else sc.closest_point(abi.REFLECTING, x)[1]
    runs = []
    for rep in range(2):
        pts = abi.HostSamplePoints(x, da, dr, nWalks=128, estimationQuantity=abi.SOLUTION, seed=0xBEEF)
        st = abi.HostStatistics(n, 3)
        steps = sc.solve(est, cases.quadratic_pde(), abi.WalkSettings.default(), pts, st)
        runs.append((steps, st.solutionMean.copy(), st.nSolutionEstimates.copy()))
    err = runs[0][1] - cases.quadratic_exact(x)
    print(name, "steps", runs[0][0], runs[1][0], "points with n != 128:", int((runs[0][2] != 128).sum()), "min n", int(runs[0][2].min()),
          "identical runs:", bool(np.array_equal(runs[0][1], runs[1][1])), "rmse %.4f" % float(np.sqrt(np.mean(err**2))), "profile", ctx.last_solve_profile())
    bad = np.where(runs[0][2] != 128)[0]
    print("   points with lost walks:", bad.tolist(), x[bad].tolist(), "walk length sums", st.totalWalkLength[bad].tolist())
    sc.close()
