"""BASELINE.json's configs at their stated sizes on ONE GPU, through the C ABI (host buffers): one JSON line per config
with wall time, walk steps/s, and the RMSE against the manufactured exact solution.

    python scripts/run_configs.py [c1] [c2] [c4] [c5] [--scale F]

--scale F < 1 runs a fraction F of the evaluation points of each config (every other size as stated) for a quick look.
C3 is bench.py's workload and is not repeated here. C5's 100 M walks are the whole job (one GPU takes all of them here;
under torchrun bench.py-style sharding each rank takes 1/N of the points)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from zombie_b200 import abi, bvc, native, procedural as P, workloads


def solve_in_batches(ctx, scene, w, batch, dist=True):
    n = len(w.points)
    sol = np.zeros(n, np.float32)
    steps, walk_ms, t0 = 0, 0.0, time.time()
    for s in range(0, n, batch):
        x = w.points[s:s + batch]
        da = scene.closest_point(abi.ABSORBING, x, want_normal=False)[1]
        dr = scene.closest_point(abi.REFLECTING, x, want_normal=False)[1] if w.reflecting is not None else None
        pts = abi.HostSamplePoints(x, da, dr, nWalks=w.n_walks, estimationQuantity=w.quantity, streams=np.arange(s, s + len(x), dtype=np.uint64))
        st = abi.HostStatistics(len(x), w.dim)
        steps += scene.solve(w.estimator, w.pde, w.settings, pts, st)
        walk_ms += ctx.last_solve_profile()["walk_ms"]
        sol[s:s + len(x)] = st.solutionMean
    return sol, steps, walk_ms, time.time() - t0


def run_walk_config(ctx, w, batch, scale):
    if scale < 1.0:
        keep = np.random.RandomState(1).permutation(len(w.points))[:max(1, int(len(w.points)*scale))]
        w.points = np.ascontiguousarray(w.points[np.sort(keep)])
    t0 = time.time()
    scene = native.Scene(ctx, w.dim, **w.scene_kwargs())
    build_s = time.time() - t0
    sol, steps, walk_ms, wall = solve_in_batches(ctx, scene, w, batch)
    err = sol - w.exact(w.points)
    out = {"config": w.name, "points": int(len(w.points)), "walks": int(len(w.points))*int(w.n_walks), "walk_steps": int(steps),
           "scene_build_s": round(build_s, 2), "solve_wall_s": round(wall, 2), "walk_kernels_s": round(walk_ms/1e3, 2),
           "steps_per_s_e2e": steps/wall, "steps_per_s_kernels": steps/(walk_ms/1e3), "rmse_vs_exact": float(np.sqrt(np.mean(err**2))),
           "scale": scale}
    scene.close()
    return out


def run_c4(ctx, scale, subdiv=7, n_eval=10_000_000, cache=(16384, 12288, 24576)):
    """BASELINE configs[3]: Boundary Value Caching (cache-point solves + splat) and Reverse Walk on Stars on a 327,680-triangle
    bumpy icosphere (the procedural family's size next above 250 K), 10 M evaluation points. 3D Poisson, Dirichlet z <= 0 /
    Neumann z > 0, u = x^2 - y^2/2 + z^2/4 + 0.3."""
    from tests import cases
    pos, tri = P.bumpy_sphere(subdiv)
    absorbing, reflecting = P.partition_mesh(pos, tri, lambda c: c[:, 2] > 0)
    pde = cases.quadratic_pde()
    ev = bvc.ManufacturedPdeEval(u=lambda x: cases.quadratic_exact(x),
                                 grad_u=lambda x: np.stack([2.0*x[:, 0], -1.0*x[:, 1], 0.5*x[:, 2]], axis=1),
                                 lap_u=lambda x: np.full(len(x), 1.5), sigma=0.0, mu=0.0)
    n_eval = max(1024, int(n_eval*scale))
    x = P.points_in_ball(n_eval, 0.7, seed=17)
    t0 = time.time()
    scene = native.Scene(ctx, 3, **cases.scene_kwargs(3, abi.REFLECTING_NEUMANN, absorbing, reflecting))
    build_s = time.time() - t0
    backend = bvc.NativeBackend(ctx, scene)
    settings = abi.WalkSettings.default(1e-3, 1e-3, 1024, False)
    offset_a = 0.03

    def inside(p):
        r = np.linalg.norm(p, axis=1)
        d = p/np.maximum(r, 1e-30)[:, None]
        return r < P.bumpy_sphere_radius(d) - np.where(d[:, 2] < 0.0, offset_a, 0.0)
    dirs = np.random.RandomState(1).normal(size=(400000, 3))
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    volume = float(4.0*np.pi/3.0*np.mean((P.bumpy_sphere_radius(dirs) - np.where(dirs[:, 2] < 0.0, offset_a, 0.0))**3))
    sa, sr = bvc.UniformBoundarySampler(*absorbing, seed=21), bvc.UniformBoundarySampler(*reflecting, seed=22)
    sd = bvc.UniformDomainSampler(inside, [-1.2]*3, [1.2]*3, volume, seed=23)
    sa.initialize(offset_a, False)
    sr.initialize(0.0, False)
    out = {"config": "C4: 3D BVC + Reverse WoSt, bumpy icosphere %d tris, %d eval points" % (len(tri), n_eval), "scene_build_s": round(build_s, 2),
           "scale": scale}
    # ---- boundary value caching
    t0 = time.time()
    solver = bvc.BoundaryValueCachingSolver(backend, sa, sr, sd, seed=31)
    solver.generate_samples(cache[0], cache[1], cache[2], offset_a, 0.0)
    t1 = time.time()
    steps = solver.compute_sample_estimates(pde, settings, 128, 512, abi.FLT_MAX, ev)
    t2 = time.time()
    da, dr = backend.distances(x)
    t3 = time.time()
    e = abi.HostBvcEvalPoints(x, da, dr)
    pairs = solver.splat(pde, e, 1e-3, 0.0, abi.FLT_MAX, 0.05, 0.05)
    t4 = time.time()
    splat_ms = ctx.last_splat_ms()
    near_steps = solver.estimate_solution_near_boundary(pde, settings, e, 0.05, 0.05, 64)
    t5 = time.time()
    err = e.solution - cases.quadratic_exact(x)
    out["bvc"] = {"cache_samples": int(solver.cache().n), "sample_generation_s": round(t1 - t0, 2), "cache_walk_steps": int(steps),
                  "cache_solve_s": round(t2 - t1, 2), "eval_distances_s": round(t3 - t2, 2), "splat_pairs": int(pairs),
                  "splat_wall_s": round(t4 - t3, 2), "splat_kernel_ms": splat_ms, "pairs_per_s_kernel": pairs/(splat_ms/1e3) if splat_ms else None,
                  "near_boundary_steps": int(near_steps), "near_boundary_s": round(t5 - t4, 2), "total_s": round(t5 - t0, 2),
                  "rmse_vs_exact": float(np.sqrt(np.mean(err**2))), "mean_error": float(err.mean())}
    # ---- reverse walk on stars onto the same evaluation points
    t0 = time.time()
    sa2, sr2 = bvc.UniformBoundarySampler(*absorbing, seed=41), bvc.UniformBoundarySampler(*reflecting, seed=42)
    sd2 = bvc.UniformDomainSampler(lambda p: P.bumpy_sphere_inside(p), [-1.2]*3, [1.2]*3, P.bumpy_sphere_volume(), seed=43)
    off = 0.01
    sa2.initialize(off, False)
    sr2.initialize(0.0, False)
    s = bvc.Samples.concatenate([sa2.generate_samples(cache[0], abi.ON_ABSORBING_BOUNDARY, off),
                                 sr2.generate_samples(cache[1], abi.ON_REFLECTING_BOUNDARY, 0.0), sd2.generate_samples(cache[2])])
    gda, gdr = backend.distances(s.pt)
    smp = abi.HostRwsSamples(s.pt, s.normal, s.type, s.pdf, gda, gdr, seed=0xBEEF)
    evr = abi.HostRwsEvalPoints(x)
    t1 = time.time()
    rsteps, records = scene.rws_solve(pde, settings, smp, evr, off)
    t2 = time.time()
    rerr = evr.estimated_solution(smp.group_counts()) - cases.quadratic_exact(x)
    out["rws"] = {"samples": int(smp.n), "walk_steps": int(rsteps), "splat_records": int(records), "sample_generation_s": round(t1 - t0, 2),
                  "solve_wall_s": round(t2 - t1, 2), "record_eval_pairs_upper_bound": int(records)*int(n_eval),
                  "rmse_vs_exact": float(np.sqrt(np.mean(rerr**2))), "mean_error": float(rerr.mean())}
    scene.close()
    return out


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    scale = 1.0
    if "--scale" in sys.argv:
        scale = float(sys.argv[sys.argv.index("--scale") + 1])
        args = [a for a in args if a != sys.argv[sys.argv.index("--scale") + 1]]
    which = args or ["c1", "c2", "c4", "c5"]
    ctx = native.Context(0)
    for c in which:
        t0 = time.time()
        if c == "c1":
            out = run_walk_config(ctx, workloads.config1(), 1 << 20, scale)
        elif c == "c2":
            out = run_walk_config(ctx, workloads.config2(), 16384, scale)
        elif c == "c5":
            out = run_walk_config(ctx, workloads.config5(subdivisions=9, n_points=1 << 20, n_walks=100), 1 << 18, scale)
        elif c == "c5mixed":
            out = run_walk_config(ctx, workloads.config5(subdivisions=9, n_points=1 << 20, n_walks=100, mixed=True), 1 << 16, scale)
        elif c == "c4":
            out = run_c4(ctx, scale)
        else:
            raise SystemExit("unknown config " + c)
        out["total_wall_s"] = round(time.time() - t0, 2)
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
