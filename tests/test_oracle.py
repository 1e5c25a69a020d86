"""CPU tests (-m "not gpu") that pin the ORACLE: the compiled reference (oracle/_ref) against the reference's
own known-answer vector (pcg32-demo.out), against the committed golden fixtures, against FCPW's own
differential test strategy (BVH vs brute-force Baseline, deps/fcpw/tests/aggregate_tests.cpp:369-622 with its
tolerances) and against exact solutions of manufactured PDEs."""
import json
import os

import numpy as np
import pytest

from tests import cases
from tests.conftest import GOLDEN, rel_err
from zombie_b200 import abi, procedural as P


def test_pcg32_known_answer_vector(ref):
    g = json.load(open(os.path.join(GOLDEN, "pcg32_demo.json")))
    u, coins, rolls, cards = ref.pcg32_demo(g["initstate"], g["initseq"], len(g["rounds"]))
    for r, rd in enumerate(g["rounds"]):
        assert [int(v) for v in u[r]] == rd["u32"]
        assert [int(v) for v in coins[r]] == rd["coins"]
        assert [int(v) for v in rolls[r]] == rd["rolls"]
        assert [int(v) for v in cards[r]] == rd["cards"]
    # SURVEY.md §8(a) row R
    assert [hex(int(v)) for v in ref.pcg32(42, 54, 6)] == ["0xa15c02b7", "0x7b47f409", "0xba1d3330", "0x83d2f293", "0xbfa4784b", "0xcbed606e"]


def test_host_pcg32_seed_matches_reference(ref):
    """zombie_b200.procedural.pcg32_seed (how the harness derives explicit per-point (state, inc))"""
    from zombie_b200.procedural import pcg32_seed
    state, inc = pcg32_seed(np.array([42, 0xC0FFEE], np.uint64), np.array([54, 7], np.uint64))
    # continue the stream in numpy and compare with the reference generator
    for k, (s0, q0) in enumerate(((42, 54), (0xC0FFEE, 7))):
        s, i = int(state[k]), int(inc[k])
        outs = []
        for _ in range(8):
            old = s
            s = (old*0x5851f42d4c957f2d + i) & 0xFFFFFFFFFFFFFFFF
            x = (((old >> 18) ^ old) >> 27) & 0xFFFFFFFF
            rot = old >> 59
            outs.append(((x >> rot) | (x << ((-rot) & 31))) & 0xFFFFFFFF)
        assert outs == [int(v) for v in ref.pcg32(s0, q0, 8)]


@pytest.mark.parametrize("dim", [2, 3])
def test_oracle_reproduces_golden_queries(ref, dim):
    g = np.load(os.path.join(GOLDEN, "queries_%dd.npz" % dim))
    whole, absorbing, reflecting = cases.mesh3d(3) if dim == 3 else cases.mesh2d(512)
    nq = len(g["cp_a_prim"])
    x = cases.query_points(nq, dim, 0.97, seed=3)
    o, d, tmax = cases.rays(nq, dim, 0.6, seed=5)
    rs = ref.RefScene(dim, **cases.scene_kwargs(dim, abi.REFLECTING_NEUMANN, absorbing, reflecting))
    prim, dist, pt, nrm = rs.closest_point(abi.ABSORBING, x)
    assert np.array_equal(prim, g["cp_a_prim"]) and np.array_equal(dist, g["cp_a_dist"]) and np.array_equal(pt, g["cp_a_pt"])
    prim, dist, pt, nrm = rs.intersect_ray(abi.REFLECTING, o, d, tmax)
    assert np.array_equal(prim, g["ray_r_prim"]) and np.array_equal(dist, g["ray_r_dist"])
    assert np.array_equal(rs.occluded(abi.ABSORBING, o, o + d*tmax[:, None]), g["occ_a"])


@pytest.mark.parametrize("dim", [2, 3])
def test_oracle_bvh_matches_bruteforce_baseline(ref, dim):
    """FCPW's own test: Bvh vs Baseline, |dd| < 1e-6 (closest point, ray), equal occlusion"""
    whole, absorbing, reflecting = cases.mesh3d(3) if dim == 3 else cases.mesh2d(512)
    kw = cases.scene_kwargs(dim, abi.REFLECTING_NEUMANN, absorbing, reflecting)
    bvh = ref.RefScene(dim, **kw)
    brute = ref.RefScene(dim, use_baseline=True, **kw)
    x = cases.query_points(512, dim, 0.97, seed=21)
    o, d, tmax = cases.rays(512, dim, 0.6, seed=23)
    for which in (abi.ABSORBING, abi.REFLECTING):
        p0, d0, pt0, _ = bvh.closest_point(which, x)
        p1, d1, pt1, _ = brute.closest_point(which, x)
        assert np.abs(d0 - d1).max() < 1e-6
        assert ((pt0 - pt1)**2).sum(axis=1).max() < 1e-6
        p0, t0, pt0, _ = bvh.intersect_ray(which, o, d, tmax)
        p1, t1, pt1, _ = brute.intersect_ray(which, o, d, tmax)
        assert np.array_equal(p0 >= 0, p1 >= 0)
        hit = p0 >= 0
        assert np.abs(t0 - t1)[hit].max() < 1e-6
        assert np.array_equal(bvh.occluded(which, o, o + d*tmax[:, None]), brute.occluded(which, o, o + d*tmax[:, None]))


def test_oracle_closest_point_against_numpy_bruteforce(ref):
    """independent restatement (Ericson closest point on triangle, vectorised) vs the compiled reference"""
    (pos, tri), _, _ = cases.mesh3d(2)
    rs = ref.RefScene(3, absorbing=(pos, tri))
    x = cases.query_points(64, 3, 1.3, seed=31)
    prim, dist, pt, _ = rs.closest_point(abi.ABSORBING, x)
    a, b, c = (pos[tri[:, k]].astype(np.float64) for k in range(3))
    best = np.full(len(x), np.inf)
    for i, q in enumerate(x.astype(np.float64)):
        best[i] = np.sqrt(_closest_dist2(a, b, c, q).min())
    assert np.abs(best - dist).max() < 2e-6
    # the returned point lies at the returned distance and on the returned triangle's plane
    assert np.abs(np.linalg.norm(x - pt, axis=1) - dist).max() < 1e-6
    n = np.cross(b - a, c - a)
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    assert np.abs(((pt - a[prim])*n[prim]).sum(axis=1)).max() < 1e-5


def _closest_dist2(a, b, c, p):
    """squared distance from p to every triangle (a[i], b[i], c[i]) — Ericson 5.1.5 in barycentric form"""
    ab, ac, ap = b - a, c - a, p - a
    d1, d2 = (ab*ap).sum(1), (ac*ap).sum(1)
    bp = p - b
    d3, d4 = (ab*bp).sum(1), (ac*bp).sum(1)
    cp = p - c
    d5, d6 = (ab*cp).sum(1), (ac*cp).sum(1)
    vc = d1*d4 - d3*d2
    vb = d5*d2 - d1*d6
    va = d3*d6 - d5*d4
    out = np.empty(len(a))
    done = np.zeros(len(a), bool)

    def put(mask, pts):
        m = mask & ~done
        out[m] = ((p - pts[m])**2).sum(1)
        done[m] = True
    put((d1 <= 0) & (d2 <= 0), a)
    put((d3 >= 0) & (d4 <= d3), b)
    put((d6 >= 0) & (d5 <= d6), c)
    with np.errstate(divide="ignore", invalid="ignore"):
        put((vc <= 0) & (d1 >= 0) & (d3 <= 0), a + ab*(d1/(d1 - d3))[:, None])
        put((vb <= 0) & (d2 >= 0) & (d6 <= 0), a + ac*(d2/(d2 - d6))[:, None])
        put((va <= 0) & (d4 - d3 >= 0) & (d5 - d6 >= 0), b + (c - b)*((d4 - d3)/((d4 - d3) + (d5 - d6)))[:, None])
        den = 1.0/(va + vb + vc)
        put(np.ones(len(a), bool), a + ab*(vb*den)[:, None] + ac*(vc*den)[:, None])
    return out


def test_oracle_golden_solves_are_reproducible_and_converge(ref):
    """the reference solver with explicit per-point pcg32 streams is deterministic (any TBB schedule) and
    its estimates match the exact solution of the manufactured problem within their own standard error"""
    g = np.load(os.path.join(GOLDEN, "solve_wos3d.npz"))
    (pos, tri), _, _ = cases.mesh3d(3)
    rs = ref.RefScene(3, absorbing=(pos, tri))
    n = 96
    pts = abi.HostSamplePoints(g["pt"][:n], g["distA"][:n], None, nWalks=128, estimationQuantity=abi.SOLUTION)
    st = abi.HostStatistics(n, 3)
    rs.solve(abi.WALK_ON_SPHERES, cases.quadratic_pde(), abi.WalkSettings.default(), pts, st, threads=2)
    assert np.array_equal(st.solutionMean, g["mean"][:n])
    assert np.array_equal(st.totalWalkLength, g["walkLength"][:n])
    se = np.sqrt(st.solution_variance()/st.nSolutionEstimates)
    z = (st.solutionMean - cases.quadratic_exact(g["pt"][:n]))/np.maximum(se, 1e-6)
    assert np.mean(np.abs(z) < 3.5) > 0.97 and abs(z.mean()) < 0.5


@pytest.mark.parametrize("name,exact", [("wost3d_neumann", "quad"), ("wost3d_robin_grad", "exp"), ("wos2d", "quad"),
                                        ("wost2d_neumann_window", "quad")])
def test_golden_solves_match_exact_solution(name, exact):
    """the committed reference estimates agree with the analytic solution (sign conventions of the
    manufactured PDE families, SURVEY.md §8c "Exact solutions")"""
    g = np.load(os.path.join(GOLDEN, "solve_%s.npz" % name))
    x = g["pt"]
    if exact == "quad":
        u = cases.quadratic_exact(x)
    else:
        _, k = cases.exp_pde(1.0, 1.0)
        u = np.exp(x @ k)
    n = np.maximum(1, g["n"])
    se = np.sqrt(g["m2"]/np.maximum(1, n - 1)/n)
    z = (g["mean"] - u)/np.maximum(se, 1e-6)
    assert np.mean(np.abs(z) < 4.0) > 0.97, (np.mean(np.abs(z) < 4.0), z.mean())
    assert abs(z.mean()) < 0.6
    if name == "wost3d_robin_grad":
        seg = np.sqrt(g["gm2"]/np.maximum(1, g["ngrad"] - 1)[:, None]/np.maximum(1, g["ngrad"])[:, None])
        zg = (g["gmean"] - u[:, None]*k[None, :])/np.maximum(seg, 1e-6)
        assert np.mean(np.abs(zg) < 4.0) > 0.97
