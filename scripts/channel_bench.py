"""Cost of extra value channels (zb_solve_channels) on the headline workload: walk steps/s of a scalar solve and of 3- and
8-channel solves on the same points. Usage: python scripts/channel_bench.py [subdivisions=8] [points=16384] [walks=64]"""
import ctypes as C
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from zombie_b200 import abi, native, workloads  # noqa: E402

sub, npts, walks = (int(sys.argv[i]) if len(sys.argv) > i else d for i, d in ((1, 8), (2, 16384), (3, 64)))
w = workloads.config3(subdivisions=sub, n_points=npts, n_walks=walks)
ctx = native.Context(0)
scene = native.Scene(ctx, 3, **w.scene_kwargs())
_, da, _, _ = scene.closest_point(abi.ABSORBING, w.points)
_, dr, _, _ = scene.closest_point(abi.REFLECTING, w.points)


def channel(k):
    d = abi.PdeDesc()
    C.memmove(C.byref(d), C.byref(w.pde), C.sizeof(abi.PdeDesc))
    if k:   # another direction of the same exponential family
        v = np.array([np.cos(k), np.sin(k), 0.3*k]); v /= np.linalg.norm(v)
        for i in range(3):
            d.params[i] = float(v[i]*np.sqrt(w.pde.absorptionCoeff))
    return d


out = {}
for c in (1, 3, 8, 1):
    pdes = [channel(k) for k in range(c)]
    best = 0.0
    for rep in range(2):
        pts = abi.HostSamplePoints(w.points, da, dr, nWalks=w.n_walks, estimationQuantity=w.quantity)
        stats = [abi.HostStatistics(pts.n, 3) for _ in range(c)]
        t = time.perf_counter()
        steps = scene.solve_channels(w.estimator, pdes, w.settings, pts, stats)
        best = max(best, steps/(time.perf_counter() - t))
    out.setdefault("channels_%d" % c, []).append(round(best/1e6, 3))
print(json.dumps({"workload": w.name, "points": npts, "walks": walks, "Msteps_per_s": out}))
