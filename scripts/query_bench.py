#!/usr/bin/env python
"""Micro-benchmark of the two cooperative queries on the headline scene (C3): zb_closest_point on the evaluation
points and zb_star_radius with the walk's search radius (distance to the absorbing boundary). Host-buffer entry
points, wall clock (the copies are < 1 % at these sizes). Used under ncu to profile the cooperative traversal in
isolation: python scripts/query_bench.py [n_queries] [subdiv]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zombie_b200 import abi, native, workloads  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
    subdiv = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    w = workloads.config3(subdiv, n, 64)
    ctx = native.Context(0)
    scene = native.Scene(ctx, 3, **w.scene_kwargs())
    x = w.points
    scene.closest_point(abi.ABSORBING, x[:1024], want_normal=False)  # warm up
    for which, name in ((abi.ABSORBING, "absorbing"), (abi.REFLECTING, "reflecting")):
        t0 = time.perf_counter()
        _, d, _, _ = scene.closest_point(which, x, want_normal=False)
        dt = time.perf_counter() - t0
        print("closest point (%s): %d queries in %.1f ms = %.1f M queries/s" % (name, n, dt*1e3, n/dt/1e6))
        if which == abi.ABSORBING:
            da = d
    minr = np.full(n, 1e-3, np.float32)
    scene.star_radius(x[:1024], minr[:1024], da[:1024])
    t0 = time.perf_counter()
    r = scene.star_radius(x, minr, da)
    dt = time.perf_counter() - t0
    print("star radius: %d queries in %.1f ms = %.2f M queries/s (%.1f %% shrunk)" % (n, dt*1e3, n/dt/1e6, 100.0*np.mean(r < 0.999*da)))
    # walk-like queries: points close to the Robin boundary, where star queries are heavy
    y = x/np.linalg.norm(x, axis=1, keepdims=True)*0.8
    y[:, 2] = np.abs(y[:, 2])
    _, dy, _, _ = scene.closest_point(abi.ABSORBING, y, want_normal=False)
    t0 = time.perf_counter()
    r = scene.star_radius(y, minr, dy)
    dt = time.perf_counter() - t0
    print("star radius near the Robin boundary: %d queries in %.1f ms = %.2f M queries/s" % (n, dt*1e3, n/dt/1e6))


if __name__ == "__main__":
    main()
