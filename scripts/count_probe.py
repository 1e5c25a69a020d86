"""Traversal-counter probe (needs the -DZB_STATS build: ZOMBIE_B200_LIB=zombie_b200/libzombie_b200_stats.so)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from zombie_b200 import abi, native, procedural as P

def counters(ctx):
    out = (C.c_ulonglong*16)()
    native.lib().zb_debug_counters(ctx._h, out)
    return list(out)

def main():
    subdiv = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    ctx = native.Context(0)
    p, t = P.bumpy_sphere(subdiv)
    (ap, ai), (rp, ri) = P.partition_mesh(p, t, lambda c: c[:, 2] > 0)
    x = P.points_in_ball(n, 0.7, seed=13)
    names = ["closest", "ray", "occlusion", "sample", "star"]
    for kind, mu, nm in ((abi.REFLECTING_NEUMANN, 0.0, "neumann"), (abi.REFLECTING_ROBIN, 1.0, "robin")):
        mn = np.full(len(ri), mu, np.float32)
        sc = native.Scene(ctx, 3, (ap, ai), (rp, ri), kind, mn, mn)
        _, da, _, _ = sc.closest_point(0, x)
        _, dr, _, _ = sc.closest_point(1, x)
        counters(ctx)
        pde = abi.PdeDesc.manufactured(abi.SOL_QUADRATIC, [1, -.5, .25, .3], robinCoeff=mu)
        pts = abi.HostSamplePoints(x, da, dr, nWalks=16, estimationQuantity=abi.SOLUTION)
        st = abi.HostStatistics(n, 3)
        steps = sc.solve(abi.WALK_ON_STARS, pde, abi.WalkSettings.default(), pts, st)
        c = counters(ctx)
        print(nm, "tris", len(ri), "steps", steps, "walk ms", ctx.last_solve_profile()["walk_ms"])
        for q, name in enumerate(names):
            nq = max(1, c[10 + q])
            print("   %-10s queries/step %.3f  nodes/query %8.2f  prims/query %8.2f" % (name, c[10 + q]/steps, c[2*q]/nq, c[2*q + 1]/nq))
        # random-point star queries for comparison
        r = sc.star_radius(x, np.full(n, 1e-3, np.float32), da)
        c = counters(ctx)
        print("   star (random interior points): nodes/query %.2f prims/query %.2f" % (c[8]/max(1, c[14]), c[9]/max(1, c[14])))
        sc.close()

if __name__ == "__main__":
    main()
