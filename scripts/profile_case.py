"""One solve of a scaled-down C3 for ncu: python scripts/profile_case.py <subdiv> <points> <walks> [quantity]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from zombie_b200 import abi, native, workloads

subdiv, n, W = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
w = workloads.config3(subdiv, n, W)
if len(sys.argv) > 4:
    w.quantity = int(sys.argv[4])
ctx = native.Context(0)
scene = native.Scene(ctx, 3, **w.scene_kwargs())
_, da, _, _ = scene.closest_point(abi.ABSORBING, w.points)
_, dr, _, _ = scene.closest_point(abi.REFLECTING, w.points)
for rep in range(int(os.environ.get('ZB_REPS', '2'))):
    pts = abi.HostSamplePoints(w.points, da, dr, nWalks=W, estimationQuantity=w.quantity)
    st = abi.HostStatistics(n, 3)
    steps = scene.solve(w.estimator, w.pde, w.settings, pts, st)
    prof = ctx.last_solve_profile()
    print("steps", steps, prof, "Msteps/s %.2f" % (steps/prof["walk_ms"]/1e3), "rmse", float(np.sqrt(np.mean((st.solutionMean - w.exact(w.points))**2))))
