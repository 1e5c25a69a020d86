"""Diagnostic run on a GPU box: queries and estimators of libzombie_b200.so against the reference oracle."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from zombie_b200 import abi, native, procedural as P
from oracle import ref

def rel(a, b):
    return np.abs(a - b)/np.maximum(1e-12, np.maximum(np.abs(a), np.abs(b)))

def main():
    subdiv = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    nq = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
    ctx = native.Context(0)
    print("KAT", [hex(v) for v in ctx.pcg32(42, 54, 6)], [hex(v) for v in ref.pcg32(42, 54, 6)])
    p, t = P.bumpy_sphere(subdiv)
    (ap, ai), (rp, ri) = P.partition_mesh(p, t, lambda c: c[:, 2] > 0)
    print("mesh", t.shape, len(ai), len(ri))
    rng = np.random.RandomState(1)
    for kind, mu in ((abi.REFLECTING_ROBIN, 1.0), (abi.REFLECTING_NEUMANN, 0.0)):
        mn = np.full(len(ri), mu, np.float32)
        t0 = time.time(); rs = ref.RefScene(3, (ap, ai), (rp, ri), kind, mn, mn); t1 = time.time()
        gs = native.Scene(ctx, 3, (ap, ai), (rp, ri), kind, mn, mn); t2 = time.time()
        print("kind", kind, "build ref %.2fs gpu %.2fs" % (t1 - t0, t2 - t1), gs.stats(0), gs.stats(1))
        x = P.points_in_ball(nq, 0.95, seed=3)
        for which in (0, 1):
            t0 = time.time(); rprim, rd, rpt, rn = rs.closest_point(which, x); t1 = time.time()
            gprim, gd, gpt, gn = gs.closest_point(which, x); t2 = time.time()
            mism = rprim != gprim
            print(" closest[%d] ref %.3fs gpu %.3fs prim mismatch %d/%d  max rel d %.3g  max |dp| %.3g  mism max rel d %.3g normal err %.3g" % (
                which, t1 - t0, t2 - t1, mism.sum(), nq, rel(rd, gd).max(), np.abs(rpt - gpt)[~mism].max(),
                rel(rd, gd)[mism].max() if mism.any() else 0, np.abs(rn - gn)[~mism].max()))
        # rays
        o = P.points_in_ball(nq, 0.6, seed=5)
        d = rng.normal(size=(nq, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
        tmax = rng.uniform(0.1, 2.0, nq).astype(np.float32)
        for which in (0, 1):
            rprim, rd, rpt, rn = rs.intersect_ray(which, o, d, tmax)
            gprim, gd, gpt, gn = gs.intersect_ray(which, o, d, tmax)
            mism = rprim != gprim
            hit = rprim >= 0
            ok = hit & ~mism
            print(" ray[%d] hits %d prim mismatch %d  max rel t %.3g max |dp| %.3g" % (which, hit.sum(), mism.sum(),
                  rel(rd, gd)[ok].max() if ok.any() else 0, np.abs(rpt - gpt)[ok].max() if ok.any() else 0))
            b = o + d*tmax[:, None]
            ro = rs.occluded(which, o, b); go = gs.occluded(which, o, b)
            print(" occluded[%d] ref %d gpu %d mismatch %d" % (which, ro.sum(), go.sum(), (ro != go).sum()))
        # star radius
        xs = P.points_in_ball(nq, 0.99, seed=9)
        _, da, _, _ = rs.closest_point(0, xs)
        minr = np.full(nq, 1e-3, np.float32)
        t0 = time.time(); rr = rs.star_radius(xs, minr, da); t1 = time.time(); gr = gs.star_radius(xs, minr, da); t2 = time.time()
        bad = np.abs(rr - gr) > 1e-3*np.maximum(1, rr)
        print(" star radius ref %.3fs gpu %.3fs: max abs diff %.3g, >1e-3: %d, frac shrunk %.3f" % (t1 - t0, t2 - t1, np.abs(rr - gr).max(), bad.sum(), (rr < da*0.999).mean()))
        if bad.any():
            j = np.where(bad)[0][:5]
            print("   e.g.", xs[j], rr[j], gr[j], da[j])
        # sphere sampling: compare the marginal distribution through the mean pdf-weighted estimate of area
        xs2 = P.points_in_ball(20000, 0.9, seed=11)
        rad = np.full(len(xs2), 0.5, np.float32)
        u = rng.uniform(size=(len(xs2), 3)).astype(np.float32)
        rprim, rpt, rn, rpdf = rs.sample_sphere(xs2, rad, u)
        gprim, gpt, gn, gpdf = gs.sample_sphere(xs2, rad, u)
        def est(prim, pdf):
            w = np.where((prim >= 0) & (pdf > 0), 1.0/np.maximum(pdf, 1e-30), 0.0)
            return w.mean(), w.std()/np.sqrt(len(w))
        print(" sample sphere: found ref %d gpu %d; area estimate ref %.4f+-%.4f gpu %.4f+-%.4f" % ((rprim >= 0).sum(), (gprim >= 0).sum(), *est(rprim, rpdf), *est(gprim, gpdf)))
        d2 = np.linalg.norm(gpt - xs2, axis=1)
        print("   gpu samples inside sphere: %.4f (ref %.4f)" % ((d2[gprim >= 0] <= 0.5 + 1e-5).mean(), (np.linalg.norm(rpt - xs2, axis=1)[rprim >= 0] <= 0.5 + 1e-5).mean()))

        # estimator
        npts = 2048
        xe = P.points_in_ball(npts, 0.7, seed=13)
        _, da, _, _ = rs.closest_point(0, xe)
        _, dr, _, _ = rs.closest_point(1, xe)
        sigma = 1.0 if kind == abi.REFLECTING_ROBIN else 0.0
        k = np.array([1, 1, 1])/3**.5 if sigma > 0 else None
        for est_kind, quantity in ((abi.WALK_ON_STARS, abi.SOLUTION), (abi.WALK_ON_STARS, abi.SOLUTION_AND_GRADIENT)):
            if sigma > 0:
                pde = abi.PdeDesc.manufactured(abi.SOL_EXP_DIR, list(k) + [1.0], absorptionCoeff=sigma, robinCoeff=mu)
                uex = np.exp(xe @ k); gex = uex[:, None]*k[None, :]
            else:
                pde = abi.PdeDesc.manufactured(abi.SOL_QUADRATIC, [1.0, -0.5, 0.25, 0.3], absorptionCoeff=0.0, robinCoeff=mu)
                uex = xe[:, 0]**2 - 0.5*xe[:, 1]**2 + 0.25*xe[:, 2]**2 + 0.3
                gex = np.stack([2*xe[:, 0], -xe[:, 1], 0.5*xe[:, 2]], 1)
            ws = abi.WalkSettings.default()
            W = 64
            rp_ = abi.HostSamplePoints(xe, da, dr, nWalks=W, estimationQuantity=quantity)
            rst = abi.HostStatistics(npts, 3)
            rsteps, rsecs = rs.solve(est_kind, pde, ws, rp_, rst)
            gp_ = abi.HostSamplePoints(xe, da, dr, nWalks=W, estimationQuantity=quantity)
            gst = abi.HostStatistics(npts, 3)
            t0 = time.time(); gsteps = gs.solve(est_kind, pde, ws, gp_, gst); gsecs = time.time() - t0
            se = np.sqrt(rst.solution_variance()/np.maximum(1, rst.nSolutionEstimates) + gst.solution_variance()/np.maximum(1, gst.nSolutionEstimates))
            z = (gst.solutionMean - rst.solutionMean)/np.maximum(se, 1e-9)
            print(" solve q=%d: ref steps %d (%.0f/s) gpu steps %d (%.0f/s) | rmse ref %.4f gpu %.4f | z mean %.3f std %.3f frac|z|<3 %.4f | meanlen ref %.2f gpu %.2f | nSol ref %d gpu %d" % (
                quantity, rsteps, rsteps/rsecs, gsteps, gsteps/gsecs, np.sqrt(np.mean((rst.solutionMean - uex)**2)), np.sqrt(np.mean((gst.solutionMean - uex)**2)),
                z.mean(), z.std(), (np.abs(z) < 3).mean(), rst.mean_walk_length().mean(), gst.mean_walk_length().mean(), rst.nSolutionEstimates.sum(), gst.nSolutionEstimates.sum()))
            if quantity == abi.SOLUTION_AND_GRADIENT:
                seg = np.sqrt(rst.gradient_variance()/np.maximum(1, rst.nGradientEstimates)[:, None] + gst.gradient_variance()/np.maximum(1, gst.nGradientEstimates)[:, None])
                zg = (gst.gradientMean - rst.gradientMean)/np.maximum(seg, 1e-9)
                print("   grad rmse ref %.4f gpu %.4f | z mean %.3f std %.3f frac<3 %.4f | firstSphere equal %s" % (
                    np.sqrt(np.mean((rst.gradientMean - gex)**2)), np.sqrt(np.mean((gst.gradientMean - gex)**2)), zg.mean(), zg.std(), (np.abs(zg) < 3).mean(),
                    np.allclose(rp_.firstSphereRadius, gp_.firstSphereRadius)))
            else:
                print("   firstSphereRadius max diff %.3g" % np.abs(rp_.firstSphereRadius - gp_.firstSphereRadius).max())
        gs.close(); rs.close()
    # WoS, dirichlet only
    rs = ref.RefScene(3, (p, t)); gs = native.Scene(ctx, 3, (p, t))
    xe = P.points_in_ball(2048, 0.7, seed=13)
    _, da, _, _ = rs.closest_point(0, xe)
    pde = abi.PdeDesc.manufactured(abi.SOL_QUADRATIC, [1.0, -0.5, 0.25, 0.3])
    uex = xe[:, 0]**2 - 0.5*xe[:, 1]**2 + 0.25*xe[:, 2]**2 + 0.3
    for quantity in (abi.SOLUTION, abi.SOLUTION_AND_GRADIENT):
        ws = abi.WalkSettings.default()
        rp_ = abi.HostSamplePoints(xe, da, None, nWalks=64, estimationQuantity=quantity); rst = abi.HostStatistics(2048, 3)
        rsteps, rsecs = rs.solve(abi.WALK_ON_SPHERES, pde, ws, rp_, rst)
        gp_ = abi.HostSamplePoints(xe, da, None, nWalks=64, estimationQuantity=quantity); gst = abi.HostStatistics(2048, 3)
        t0 = time.time(); gsteps = gs.solve(abi.WALK_ON_SPHERES, pde, ws, gp_, gst); gsecs = time.time() - t0
        se = np.sqrt(rst.solution_variance()/rst.nSolutionEstimates + gst.solution_variance()/gst.nSolutionEstimates)
        z = (gst.solutionMean - rst.solutionMean)/np.maximum(se, 1e-9)
        print("WoS q=%d: ref steps %d (%.0f/s) gpu %d (%.0f/s) rmse ref %.4f gpu %.4f z mean %.3f std %.3f frac<3 %.4f" % (quantity, rsteps, rsteps/rsecs, gsteps, gsteps/gsecs,
              np.sqrt(np.mean((rst.solutionMean - uex)**2)), np.sqrt(np.mean((gst.solutionMean - uex)**2)), z.mean(), z.std(), (np.abs(z) < 3).mean()))

if __name__ == "__main__":
    main()
