"""Timing probe: walk kernel duration versus problem size and enabled contributions."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from zombie_b200 import abi, native, procedural as P

def run(ctx, scene, x, da, dr, pde, ws, est, W, q, label):
    pts = abi.HostSamplePoints(x, da, dr, nWalks=W, estimationQuantity=q)
    st = abi.HostStatistics(len(x), 3)
    t0 = time.time()
    steps = scene.solve(est, pde, ws, pts, st)
    wall = time.time() - t0
    prof = ctx.last_solve_profile()
    print("%-44s n=%7d walks=%8d steps=%10d prepare %8.2f ms walk %9.2f ms reduce %7.2f ms wall %7.3f s -> %8.2f Msteps/s (kernel) maxlen %d grid %d" % (
        label, len(x), prof["walks"], steps, prof["prepare_ms"], prof["walk_ms"], prof["reduce_ms"], wall,
        steps/max(prof["walk_ms"], 1e-3)/1e3, int((st.totalWalkLength/np.maximum(1, st.nSolutionEstimates)).max()), prof["walk_grid"]), flush=True)

def main():
    subdiv = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    sizes = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [2048, 16384]
    ctx = native.Context(0)
    p, t = P.bumpy_sphere(subdiv)
    (ap, ai), (rp, ri) = P.partition_mesh(p, t, lambda c: c[:, 2] > 0)
    mu1 = np.ones(len(ri), np.float32)
    k = [3**-.5]*3 + [1.0]
    for n in sizes:
        x = P.points_in_ball(n, 0.7, seed=13)
        sd = native.Scene(ctx, 3, (p, t))
        _, da0, _, _ = sd.closest_point(0, x)
        run(ctx, sd, x, da0, None, abi.PdeDesc.manufactured(abi.SOL_QUADRATIC, [1, -.5, .25, .3]), abi.WalkSettings.default(), abi.WALK_ON_SPHERES, 64, abi.SOLUTION, "WoS dirichlet harmonic")
        run(ctx, sd, x, da0, None, abi.PdeDesc.manufactured(abi.SOL_QUADRATIC, [1, -.5, .25, .3]), abi.WalkSettings.default(), abi.WALK_ON_STARS, 64, abi.SOLUTION, "WoSt no reflecting boundary, harmonic")
        sd.close()
        for kind, mu, sigma, name in ((abi.REFLECTING_NEUMANN, 0.0, 0.0, "neumann harmonic"), (abi.REFLECTING_ROBIN, 1.0, 0.0, "robin harmonic"), (abi.REFLECTING_ROBIN, 1.0, 1.0, "robin yukawa")):
            mn = np.full(len(ri), mu, np.float32)
            sc = native.Scene(ctx, 3, (ap, ai), (rp, ri), kind, mn, mn)
            _, da, _, _ = sc.closest_point(0, x)
            _, dr, _, _ = sc.closest_point(1, x)
            if sigma > 0:
                pde = abi.PdeDesc.manufactured(abi.SOL_EXP_DIR, k, absorptionCoeff=sigma, robinCoeff=mu)
            else:
                pde = abi.PdeDesc.manufactured(abi.SOL_QUADRATIC, [1, -.5, .25, .3], robinCoeff=mu)
            run(ctx, sc, x, da, dr, pde, abi.WalkSettings.default(), abi.WALK_ON_STARS, 64, abi.SOLUTION, "WoSt " + name)
            run(ctx, sc, x, da, dr, pde, abi.WalkSettings.default(ignoreReflectingBoundaryContribution=1), abi.WALK_ON_STARS, 64, abi.SOLUTION, "WoSt " + name + " -refl")
            run(ctx, sc, x, da, dr, pde, abi.WalkSettings.default(ignoreReflectingBoundaryContribution=1, ignoreSourceContribution=1), abi.WALK_ON_STARS, 64, abi.SOLUTION, "WoSt " + name + " -refl -src")
            run(ctx, sc, x, da, dr, pde, abi.WalkSettings.default(maxWalkLength=64), abi.WALK_ON_STARS, 64, abi.SOLUTION, "WoSt " + name + " maxlen64")
            run(ctx, sc, x, da, dr, pde, abi.WalkSettings.default(), abi.WALK_ON_STARS, 64, abi.SOLUTION_AND_GRADIENT, "WoSt " + name + " +gradient")
            sc.close()

if __name__ == "__main__":
    main()
