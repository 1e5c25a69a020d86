"""Boundary Value Caching splat (BASELINE config 4): zb_bvc_splat against the reference's
bvc::BoundaryValueCaching::splat (zombie/variance_reduction/boundary_value_caching.h:371-502, 682-811).

Deterministic floating-point arithmetic, so the gate is numeric: per evaluation point the five bucket counts are
equal and every bucket mean / the summed solution and gradient agree within RTOL of the magnitude of the
quantity (float32 sums of ~10^3 terms in a different order than the reference's running Welford mean)."""
import math

import numpy as np
import pytest

from zombie_b200 import abi, native, procedural as P

RTOL = 2e-4


def make_problem(dim, n_samples, n_eval, seed, on_boundary_eval=False):
    """cache samples on / inside the unit sphere (circle) with every sample kind the splat distinguishes"""
    rng = np.random.RandomState(seed)
    y = rng.normal(size=(n_samples, dim))
    y /= np.linalg.norm(y, axis=1, keepdims=True)
    normal = y.copy()
    typ = rng.choice([abi.IN_DOMAIN, abi.ON_ABSORBING_BOUNDARY, abi.ON_REFLECTING_BOUNDARY], n_samples, p=[0.3, 0.3, 0.4])
    y[typ == abi.IN_DOMAIN] *= rng.uniform(0.0, 1.0, (int((typ == abi.IN_DOMAIN).sum()), 1))**(1.0/dim)
    aligned = (rng.rand(n_samples) < 0.25) & (typ != abi.IN_DOMAIN)
    robin = np.where(typ == abi.ON_REFLECTING_BOUNDARY, rng.choice([0.0, 0.4, 3.0], n_samples), 0.0)
    samples = abi.HostBvcSamples(y, normal, typ, rng.uniform(0.05, 0.3, n_samples), rng.normal(size=n_samples),
                                 rng.normal(size=n_samples), rng.normal(size=n_samples), robin, aligned)
    x = P.points_in_ball(n_eval, 0.9, seed=seed + 1, dim=dim)
    dist = (1.0 - np.linalg.norm(x, axis=1)).astype(np.float32)
    etype = np.zeros(n_eval, np.int32)
    if on_boundary_eval:
        etype[::7] = abi.ON_REFLECTING_BOUNDARY
    ev = abi.HostBvcEvalPoints(x, dist, dist + 0.01, etype)
    return samples, ev


def numpy_splat(dim, sigma, s, e, radius_clamp, reg, robin_cutoff, cutoff_a, cutoff_r):
    """float32 restatement of splatBoundaryData / splatSourceData with plain sums (what the kernel computes)"""
    f = np.float32
    n, m = e.n, s.n
    xy = e.pt[:, None, :] - s.pt[None, :, :]
    sign = np.where(s.estimateBoundaryNormalAligned != 0, f(-1), f(1))
    nrm = s.normal*sign[:, None]
    r = np.maximum(f(radius_clamp), np.sqrt((xy*xy).sum(-1, dtype=f)))
    ndot = (xy*nrm[None]).sum(-1, dtype=f)
    if dim == 3:
        c = f(1.0/(4.0*math.pi))
        if sigma > 0:
            mu = f(math.sqrt(sigma))
            mur = r*mu
            ex = np.exp(-mur)
            q1 = mu*ex*(1 + 1/mur)
            q2 = mu*ex/mur
            G = ex*c/r
            Pk = ndot*q1*c/(r*r)
            dG = -xy*(q1*c/(r*r))[..., None]
            dP = (nrm[None]*q1[..., None] - ((ndot/(r*r))*(2*q1 + q2))[..., None]*xy)*(c/(r*r))[..., None]
        else:
            G = c/r
            Pk = ndot*c/r**3
            dG = -xy*(c/r**3)[..., None]
            dP = (nrm[None] - (3*ndot/(r*r))[..., None]*xy)*(c/r**3)[..., None]
    elif sigma > 0:
        from scipy.special import kv
        c = f(1.0/(2.0*math.pi))
        mu = f(math.sqrt(sigma))
        mur = (r*mu).astype(np.float64)
        K0, K1, K2 = kv(0, mur).astype(f), kv(1, mur).astype(f), kv(2, mur).astype(f)
        q1, q2 = mu*K1, f(sigma)*(K0 + K2)/2
        G = K0*c
        Pk = ndot*q1*c/r
        dG = -xy*(q1*c/r)[..., None]
        dP = (nrm[None]*q1[..., None] - ((ndot/(r*r))*(q1 + r*q2))[..., None]*xy)*(c/r)[..., None]
    else:
        c = f(1.0/(2.0*math.pi))
        G = -np.log(r)*c
        Pk = ndot*c/(r*r)
        dG = -xy*(c/(r*r))[..., None]
        dP = (nrm[None] - (2*ndot/(r*r))[..., None]*xy)*(c/(r*r))[..., None]
    if reg > 0:
        rr = r/f(reg)
        if dim == 3:
            from scipy.special import erf
            G = G*erf(rr)
            Pk = Pk*(erf(rr) - 2*rr*np.exp(-rr*rr)/math.sqrt(math.pi))
        else:
            Pk = Pk*(1 - np.exp(-rr*rr))
    sol, nd, rob, mu_s, pdf = s.solution[None], s.normalDerivative[None], s.contribution[None], s.robinCoeff[None], s.pdf[None]
    src = (s.type == abi.IN_DOMAIN)[None]
    case_a = np.abs(mu_s) > robin_cutoff
    case_b = (~case_a) & (np.abs(mu_s) > 0)
    with np.errstate(divide="ignore", invalid="ignore"):
        ua = (G + Pk/mu_s)*nd - Pk*rob/mu_s
        ga = (dG + dP/mu_s[..., None])*nd[..., None] - dP*(rob/mu_s)[..., None]
    ub = G*rob - (Pk + mu_s*G)*sol
    gb = dG*rob[..., None] - (dP + mu_s[..., None]*dG)*sol[..., None]
    uc = G*nd - Pk*sol
    gc = dG*nd[..., None] - dP*sol[..., None]
    u = np.where(src, G*rob, np.where(case_a, ua, np.where(case_b, ub, uc)))
    g = np.where(src[..., None], dG*rob[..., None], np.where(case_a[..., None], ga, np.where(case_b[..., None], gb, gc)))
    on_b = (e.type != abi.IN_DOMAIN)
    alpha = np.where(on_b, f(2), f(1))[:, None]
    alpha_g = np.where(on_b, f(0), f(1))[:, None]
    u = alpha*u/pdf
    g = alpha_g[..., None]*g/pdf[..., None]
    bucket = np.where(s.type == abi.ON_ABSORBING_BOUNDARY, 0, np.where(s.type == abi.ON_REFLECTING_BOUNDARY, 2, 4))
    bucket = bucket + ((s.estimateBoundaryNormalAligned != 0) & (bucket < 4))
    live = ~((e.distToAbsorbingBoundary < cutoff_a) | (e.distToReflectingBoundary < cutoff_r))
    cnt = np.zeros((n, 5), np.int32)
    mean = np.zeros((n, 5, 1 + dim), np.float64)
    for b in range(5):
        sel = bucket == b
        k = int(sel.sum())
        if k == 0:
            continue
        cnt[live, b] = k
        mean[live, b, 0] = u[live][:, sel].sum(1, dtype=np.float64)/k
        mean[live, b, 1:] = g[live][:, sel].sum(1, dtype=np.float64)/k
    return mean[:, :, 0].sum(1), mean[:, :, 1:].sum(1), cnt, mean


def scale_of(mean):
    """magnitude against which differences are judged: the largest bucket mean of the batch, per quantity"""
    return np.abs(mean).max(axis=(0, 1), keepdims=True) + 1e-12


CASES = [
    dict(dim=3, sigma=0.0, reg=0.0, robin_cutoff=1.0),
    dict(dim=3, sigma=2.0, reg=0.0, robin_cutoff=1.0),
    dict(dim=3, sigma=0.0, reg=0.05, robin_cutoff=abi.FLT_MAX),
    dict(dim=2, sigma=0.0, reg=0.0, robin_cutoff=1.0),
    dict(dim=2, sigma=0.0, reg=0.05, robin_cutoff=1.0),
]
CPU_ONLY_CASES = [dict(dim=2, sigma=3.0, reg=0.0, robin_cutoff=1.0)]   # its GPU run lives in tests/test_rws.py


@pytest.mark.parametrize("c", CASES + CPU_ONLY_CASES, ids=lambda c: "dim%d_sigma%g_reg%g" % (c["dim"], c["sigma"], c["reg"]))
def test_reference_splat_matches_formula_restatement(ref, c):
    """CPU: pins the oracle wrapper (oracle/ref_capi.cpp: zref_bvc_splat) and the restated formulas the kernel
    implements against each other, including the cutoff-distance rule and boundary evaluation points."""
    s, e = make_problem(c["dim"], 700, 120, seed=11, on_boundary_eval=True)
    sol, grad, cnt, mean, _ = ref.bvc_splat(c["dim"], c["sigma"], s, e, 1e-3, c["reg"], c["robin_cutoff"], 0.15, 0.0)
    nsol, ngrad, ncnt, nmean = numpy_splat(c["dim"], c["sigma"], s, e, 1e-3, c["reg"], c["robin_cutoff"], 0.15, 0.0)
    assert np.array_equal(cnt, ncnt)
    assert (cnt.sum(1) == 0).any() and (cnt.sum(1) > 0).any()  # the cutoff removed some evaluation points, not all
    assert np.all(np.abs(mean - nmean) <= RTOL*scale_of(nmean))
    assert np.all(np.abs(sol - nsol) <= 5*RTOL*scale_of(nmean)[0, 0, 0])
    assert np.all(np.abs(grad - ngrad) <= 5*RTOL*scale_of(nmean)[0, 0, 1:].max())


@pytest.mark.gpu
@pytest.mark.parametrize("c", CASES, ids=lambda c: "dim%d_sigma%g_reg%g" % (c["dim"], c["sigma"], c["reg"]))
def test_gpu_splat_matches_reference(ctx, ref, c):
    check_gpu_splat(ctx, ref, c)


def check_gpu_splat(ctx, ref, c):
    s, e = make_problem(c["dim"], 3000, 5000, seed=23, on_boundary_eval=True)
    sol, grad, cnt, mean, _ = ref.bvc_splat(c["dim"], c["sigma"], s, e, 1e-3, c["reg"], c["robin_cutoff"], 0.12, 0.0)
    pairs = ctx.bvc_splat(c["dim"], c["sigma"], s, e, 1e-3, c["reg"], c["robin_cutoff"], 0.12, 0.0)
    assert pairs == s.n*e.n
    assert np.array_equal(e.bucketCount, cnt)
    gm = e.bucket_mean()
    assert np.all(np.abs(gm - mean) <= RTOL*scale_of(mean))
    assert np.all(np.abs(e.solution - sol) <= 5*RTOL*scale_of(mean)[0, 0, 0])
    assert np.all(np.abs(e.gradient - grad) <= 5*RTOL*scale_of(mean)[0, 0, 1:].max())


@pytest.mark.gpu
def test_gpu_splat_accumulates_like_repeated_reference_splats(ctx, ref):
    """progressive refinement: splatting two halves of the cache in two calls == one call with the whole cache
    (the reference's statistics accumulate across splat calls, boundary_value_caching.h:482-502)"""
    s, e = make_problem(3, 2000, 3000, seed=5)
    whole = abi.HostBvcEvalPoints(e.pt, e.distToAbsorbingBoundary, e.distToReflectingBoundary, e.type)
    ctx.bvc_splat(3, 0.0, s, whole)
    idx = np.arange(s.n)
    ctx.bvc_splat(3, 0.0, s.subset(idx[:900]), e)
    ctx.bvc_splat(3, 0.0, s.subset(idx[900:]), e)
    assert np.array_equal(e.bucketCount, whole.bucketCount)
    assert np.allclose(e.bucket_mean(), whole.bucket_mean(), rtol=0, atol=RTOL*float(np.abs(whole.bucket_mean()).max()))
    sol, _, _, _, _ = ref.bvc_splat(3, 0.0, s, whole)
    assert np.all(np.abs(e.solution - sol) <= 5*RTOL*np.abs(sol).max())


@pytest.mark.gpu
def test_gpu_splat_edge_cases(ctx):
    s, e = make_problem(3, 10, 0, seed=1)
    assert ctx.bvc_splat(3, 0.0, s, e) == 0                       # no evaluation points
    s0, e0 = make_problem(3, 0, 33, seed=1)
    assert ctx.bvc_splat(3, 0.0, s0, e0) == 0                     # empty cache: estimates are zero
    assert not e0.solution.any() and not e0.bucketCount.any()
    with pytest.raises(native.ZombieError):
        s2, e2 = make_problem(2, 10, 10, seed=1)
        ctx.bvc_splat(2, -1.0, s2, e2)                             # a negative absorption coefficient is an error, not a crash
