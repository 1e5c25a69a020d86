#!/usr/bin/env python
"""Generates the golden fixtures of tests/golden/ from the REFERENCE itself. Run in a container that has
/root/reference (it cannot travel to the GPU box, the fixtures can):

    python -c 'import __graft_entry__ as g; g.build()'      # builds oracle/_ref from /root/reference
    python tests/golden/make_golden.py

  pcg32_demo.json   parsed from the reference's committed known-answer file deps/pcg32/pcg32-demo.out
                    (all 5 rounds: 32-bit outputs, coins = nextUInt(2), rolls = nextUInt(6)+1, shuffled cards)
  queries_{2d,3d}.npz   FCPW / ReflectanceBvh query results of the compiled reference on seeded inputs
                    (tests/cases.py): closest point, closest ray hit, occlusion, star radius (Neumann and
                    Robin handler), sphere-sample pdf sums
  solve_{...}.npz   WalkOnSpheres / WalkOnStars::solve estimates of the compiled reference with explicit
                    per-point pcg32 (seed, stream) — deterministic, independent of the TBB schedule
"""
import json
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref  # noqa: E402
from tests import cases  # noqa: E402
from zombie_b200 import abi  # noqa: E402

REF = os.environ.get("ZOMBIE_REFERENCE", "/root/reference")


def pcg32_demo():
    text = open(os.path.join(REF, "deps/pcg32/pcg32-demo.out")).read()
    rounds = []
    number = "A23456789TJQK"
    suit = "hcds"
    for block in re.split(r"Round \d+:\n", text)[1:]:
        u32 = [int(v, 16) for v in re.search(r"32bit:(.*)\n", block).group(1).split()]
        coins = [1 if c == "H" else 0 for c in re.search(r"Coins: (\S+)\n", block).group(1)]
        rolls = [int(v) for v in re.search(r"Rolls:(.*)\n", block).group(1).split()]
        cards_txt = block[block.index("Cards:") + 6:].split()
        cards = [number.index(c[0])*4 + suit.index(c[1]) for c in cards_txt]
        assert len(u32) == 6 and len(coins) == 65 and len(rolls) == 33 and len(cards) == 52
        rounds.append({"u32": u32, "coins": coins, "rolls": rolls, "cards": cards})
    json.dump({"source": "deps/pcg32/pcg32-demo.out (pcg32-demo.cpp: rng.seed(42u, 54u))", "initstate": 42, "initseq": 54,
               "rounds": rounds}, open(os.path.join(HERE, "pcg32_demo.json"), "w"), indent=0)
    print("pcg32_demo.json: %d rounds" % len(rounds))


def queries(dim, nq=768):
    if dim == 3:
        whole, absorbing, reflecting = cases.mesh3d(3)
    else:
        whole, absorbing, reflecting = cases.mesh2d(512)
    out = {}
    x = cases.query_points(nq, dim, 0.97, seed=3)
    o, d, tmax = cases.rays(nq, dim, 0.6, seed=5)
    xs = cases.query_points(nq, dim, 0.99, seed=9)
    for kind, name in ((abi.REFLECTING_NEUMANN, "neumann"), (abi.REFLECTING_ROBIN, "robin")):
        rs = ref.RefScene(dim, **cases.scene_kwargs(dim, kind, absorbing, reflecting, mu=1.0))
        if kind == abi.REFLECTING_NEUMANN:
            for which, w in ((abi.ABSORBING, "a"), (abi.REFLECTING, "r")):
                prim, dist, pt, nrm = rs.closest_point(which, x)
                out["cp_%s_prim" % w], out["cp_%s_dist" % w], out["cp_%s_pt" % w], out["cp_%s_normal" % w] = prim, dist, pt, nrm
                # bounded search radius: some queries find nothing
                rad = np.full(nq, 0.15, np.float32)
                prim, dist, _, _ = rs.closest_point(which, x, rad)
                out["cpr_%s_prim" % w], out["cpr_%s_dist" % w] = prim, dist
                prim, dist, pt, nrm = rs.intersect_ray(which, o, d, tmax)
                out["ray_%s_prim" % w], out["ray_%s_dist" % w], out["ray_%s_pt" % w], out["ray_%s_normal" % w] = prim, dist, pt, nrm
                out["occ_%s" % w] = rs.occluded(which, o, o + d*tmax[:, None])
        _, da, _, _ = rs.closest_point(abi.ABSORBING, xs)
        minr = np.full(nq, 1e-3, np.float32)
        out["star_%s" % name] = rs.star_radius(xs, minr, da)
        out["star_%s_flip" % name] = rs.star_radius(xs, minr, da, flip=np.ones(nq, np.uint8))
        out["star_maxr"] = da
        rs.close()
    np.savez_compressed(os.path.join(HERE, "queries_%dd.npz" % dim), **out)
    print("queries_%dd.npz: %d arrays" % (dim, len(out)))


def solve_case(name, dim, estimator, kind, pde, settings, pts_xyz, n_walks, quantity, absorbing, reflecting, mu=1.0):
    rs = ref.RefScene(dim, **cases.scene_kwargs(dim, kind, absorbing, reflecting, mu=mu))
    _, da, _, _ = rs.closest_point(abi.ABSORBING, pts_xyz)
    dr = None
    if reflecting is not None:
        _, dr, _, _ = rs.closest_point(abi.REFLECTING, pts_xyz)
    pts = abi.HostSamplePoints(pts_xyz, da, dr, nWalks=n_walks, estimationQuantity=quantity)
    st = abi.HostStatistics(pts.n, dim)
    steps, _ = rs.solve(estimator, pde, settings, pts, st, threads=0)
    np.savez_compressed(os.path.join(HERE, "solve_%s.npz" % name), pt=pts_xyz, distA=da,
                        distR=dr if dr is not None else np.zeros(0, np.float32),
                        mean=st.solutionMean, m2=st.solutionM2, n=st.nSolutionEstimates, gmean=st.gradientMean,
                        gm2=st.gradientM2, ngrad=st.nGradientEstimates, walkLength=st.totalWalkLength,
                        firstSphereRadius=pts.firstSphereRadius, steps=np.int64(steps))
    rs.close()
    print("solve_%s.npz: %d points, %d walk steps" % (name, pts.n, steps))


def solves():
    whole3, a3, r3 = cases.mesh3d(3)
    whole2, a2, r2 = cases.mesh2d(512)
    x3 = cases.query_points(4096, 3, 0.7, seed=13)
    x2 = cases.query_points(4096, 2, 0.6, seed=13)
    ws = abi.WalkSettings.default()
    solve_case("wos3d", 3, abi.WALK_ON_SPHERES, abi.REFLECTING_NONE, cases.quadratic_pde(), ws, x3, 128, abi.SOLUTION, whole3, None)
    solve_case("wos2d", 2, abi.WALK_ON_SPHERES, abi.REFLECTING_NONE, cases.quadratic_pde(dim=2), ws, x2, 128, abi.SOLUTION, whole2, None)
    solve_case("wost3d_neumann", 3, abi.WALK_ON_STARS, abi.REFLECTING_NEUMANN, cases.quadratic_pde(), ws, x3, 128, abi.SOLUTION, a3, r3)
    pde, _ = cases.exp_pde(1.0, 1.0)
    solve_case("wost3d_robin_grad", 3, abi.WALK_ON_STARS, abi.REFLECTING_ROBIN, pde, ws, x3, 128, abi.SOLUTION_AND_GRADIENT, a3, r3)
    ws2 = abi.WalkSettings.default(russianRouletteThreshold=0.99, splittingThreshold=1.5)
    solve_case("wost2d_neumann_window", 2, abi.WALK_ON_STARS, abi.REFLECTING_NEUMANN, cases.quadratic_pde(dim=2), ws2, x2, 128,
               abi.SOLUTION, a2, r2)


if __name__ == "__main__":
    pcg32_demo()
    queries(3)
    queries(2)
    solves()
