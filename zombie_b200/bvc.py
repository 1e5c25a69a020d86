"""Boundary Value Caching end to end (BASELINE configs[3]): the host-side orchestration of
bvc::BoundaryValueCachingSolver (zombie/variance_reduction/boundary_value_caching.h:814-925) and the boundary / domain
samplers it draws its cache from (variance_reduction/boundary_samplers.h:290-625, domain_samplers.h:74-108), over the
device entry points of the C ABI:

    generate_samples                cache points on the (offset) boundary and in the domain; their distances to the
                                    boundaries come from one batched zb_closest_point per mesh instead of two host
                                    callbacks per sample (boundary_samplers.h:604-619)
    compute_sample_estimates        zb_solve (WalkOnStars) at the cache points: solution, and the normal derivative as
                                    a directional-derivative estimate (boundary_value_caching.h:319-343, 567-679)
    splat                           zb_bvc_splat: every cache sample onto every evaluation point (:886-910)
    estimate_solution_near_boundary zb_solve at the evaluation points closer to a boundary than the cutoffs (:913-925)

Everything numerical happens on the device behind `backend`; this module only moves arrays. `backend` is duck-typed
(NativeBackend below wraps native.Scene / native.Context; the tests also run the same orchestration over the compiled
reference to pin it on a CPU-only box).

Sampling is stratified exactly like the reference (jittered Latin hypercube, CDF over primitive areas), but the
reference seeds its samplers from the clock (sampling.h:36-41), so caches agree in distribution, not sample for sample.
"""
import numpy as np

from . import abi

F = np.float32


def generate_stratified_samples(n, dims, rng):
    """generateStratifiedSamples<DIM> (zombie/core/sampling.h:436-458): jittered diagonal + one shuffle per dimension"""
    if n <= 0:
        return np.zeros((0, dims), F)
    s = (np.arange(n)[:, None] + rng.random_sample((n, dims)))/n
    s = np.minimum(s, 1.0 - np.finfo(F).eps)
    for j in range(dims):
        s[:, j] = s[rng.permutation(n), j]
    return s.astype(F)


class CdfTable:
    """CDFTable (sampling.h:252-315)"""

    def __init__(self, weights):
        w = np.asarray(weights, F)
        self.table = np.concatenate([[0.0], np.cumsum(w, dtype=F)]).astype(F)
        self.total = float(self.table[-1]) if len(w) else 0.0
        if len(w):
            self.table = (self.table/self.table[-1]).astype(F) if self.total > 0 else (np.arange(len(w) + 1)/len(w)).astype(F)

    def sample(self, u):
        # first index with table[i] > u, minus one, clamped (sampling.h:288-306)
        return np.clip(np.searchsorted(self.table, u, side="right") - 1, 0, len(self.table) - 2)


class Samples:
    """the fields of zombie::SamplePoint a cache uses (common.h:152-195), as arrays"""

    def __init__(self, pt, normal, type_, pdf, aligned):
        n = len(pt)
        self.pt = np.ascontiguousarray(pt, F)
        self.normal = np.ascontiguousarray(normal, F)
        self.type = np.full(n, type_, np.int32)
        self.pdf = np.full(n, pdf, F) if np.isscalar(pdf) else np.ascontiguousarray(pdf, F)
        self.aligned = np.full(n, 1 if aligned else 0, np.uint8)
        self.dist_a = np.zeros(n, F)
        self.dist_r = np.zeros(n, F)
        self.robin_coeff = np.zeros(n, F)
        self.solution = np.zeros(n, F)
        self.normal_derivative = np.zeros(n, F)
        self.contribution = np.zeros(n, F)

    def __len__(self):
        return len(self.pt)

    @staticmethod
    def concatenate(parts):
        parts = [p for p in parts if p is not None and len(p)]
        if not parts:
            return None
        out = Samples(np.concatenate([p.pt for p in parts]), np.concatenate([p.normal for p in parts]), 0, 1.0, False)
        for k in ("type", "pdf", "aligned", "dist_a", "dist_r", "robin_coeff", "solution", "normal_derivative", "contribution"):
            setattr(out, k, np.concatenate([getattr(p, k) for p in parts]))
        return out


class UniformBoundarySampler:
    """UniformLineSegmentBoundarySampler / UniformTriangleBoundarySampler (boundary_samplers.h:187-625): samples the
    boundary mesh in proportion to weight x area of the primitives displaced along the vertex normals."""

    def __init__(self, positions, indices, weights=None, weights_normal_aligned=None, inside_solve_region=None,
                 weighted_normals=False, seed=0x5EED):
        self.pos = np.ascontiguousarray(positions, F)
        self.idx = np.ascontiguousarray(indices, np.int64)
        self.dim = self.pos.shape[1]
        n = len(self.idx)
        self.w = np.ones(n, F) if weights is None else np.asarray(weights, F)
        self.w_aligned = self.w if weights_normal_aligned is None else np.asarray(weights_normal_aligned, F)
        if len(self.w) != n or len(self.w_aligned) != n or not np.all(np.isfinite(self.w)) or (self.w < 0).any():
            raise ValueError("UniformBoundarySampler: one finite non-negative weight per primitive is required")
        self.inside = inside_solve_region or (lambda p: np.ones(len(p), bool))
        self.rng = np.random.RandomState(seed)
        self.vertex_normals = self._vertex_normals(weighted_normals)
        self.cdf = self.cdf_aligned = None
        self.mass = self.mass_aligned = 0.0

    def _prim_normals(self, v):
        if self.dim == 3:
            return np.cross(v[1] - v[0], v[2] - v[0])
        s = v[1] - v[0]
        return np.stack([s[:, 1], -s[:, 0]], axis=1)

    def _vertex_normals(self, weighted):
        # computeNormals (boundary_samplers.h:257-288, 477-508): (angle-weighted) sum of incident primitive normals
        v = [self.pos[self.idx[:, k]] for k in range(self.dim)]
        n = self._prim_normals(v)
        n = n/np.maximum(np.linalg.norm(n, axis=1, keepdims=True), 1e-30)
        out = np.zeros_like(self.pos, dtype=np.float64)
        for j in range(self.dim):
            wgt = np.ones(len(n))
            if weighted and self.dim == 3:
                a, b, c = v[j], v[(j + 1) % 3], v[(j + 2) % 3]
                u, w = b - a, c - a
                u = u/np.maximum(np.linalg.norm(u, axis=1, keepdims=True), 1e-30)
                w = w/np.maximum(np.linalg.norm(w, axis=1, keepdims=True), 1e-30)
                wgt = np.arccos(np.clip((u*w).sum(1), -1.0, 1.0))
            np.add.at(out, self.idx[:, j], n*wgt[:, None])
        return (out/np.maximum(np.linalg.norm(out, axis=1, keepdims=True), 1e-30)).astype(F)

    def _measure(self, v):
        if self.dim == 3:
            return 0.5*np.linalg.norm(np.cross(v[1] - v[0], v[2] - v[0]), axis=1)
        return np.linalg.norm(v[1] - v[0], axis=1)

    def _build(self, offset, weights):
        # buildCDFTable (boundary_samplers.h:290-313, 510-538)
        v = [self.pos[self.idx[:, k]] for k in range(self.dim)]
        mid = sum(v)/self.dim
        n = self._prim_normals(v)
        n = n/np.maximum(np.linalg.norm(n, axis=1, keepdims=True), 1e-30)
        keep = self.inside((mid + offset*n).astype(F))
        moved = [self.pos[self.idx[:, k]] + offset*self.vertex_normals[self.idx[:, k]] for k in range(self.dim)]
        table = CdfTable(np.where(keep, self._measure(moved)*weights, 0.0))
        return table, table.total

    def initialize(self, normal_offset, solve_double_sided):
        """BoundarySampler::initialize (:315-326, 540-551): inward table always, outward table for double-sided solves"""
        self.cdf, self.mass = self._build(-normal_offset, self.w)
        if solve_double_sided:
            self.cdf_aligned, self.mass_aligned = self._build(normal_offset, self.w_aligned)

    def get_sample_count(self, n_total, normal_aligned):
        total = self.mass + self.mass_aligned
        if total <= 0:
            return 0
        return int(np.ceil(n_total*(self.mass_aligned if normal_aligned else self.mass)/total))

    def generate_samples(self, n, sample_type, normal_offset, normal_aligned=False):
        """generateSamples (:328-402, 553-645); distances are filled in by the caller in one batch"""
        table, mass, weights = (self.cdf_aligned, self.mass_aligned, self.w_aligned) if normal_aligned else (self.cdf, self.mass, self.w)
        offset = normal_offset if normal_aligned else -normal_offset
        if table is None or mass <= 0 or n <= 0:
            return Samples(np.zeros((0, self.dim), F), np.zeros((0, self.dim), F), sample_type, 1.0, normal_aligned)
        prim = table.sample(generate_stratified_samples(n, 1, self.rng)[:, 0])
        order = np.argsort(prim, kind="stable")
        prim = prim[order]
        # stratified samples inside every primitive that was hit more than once (:580-590)
        u = np.zeros((n, max(1, self.dim - 1)), F)
        start = 0
        uniq, counts = np.unique(prim, return_counts=True)
        for k, c in zip(uniq, counts):
            u[start:start + c] = (self.rng.random_sample((1, self.dim - 1)) if c == 1
                                  else generate_stratified_samples(c, self.dim - 1, self.rng))
            start += c
        v = [self.pos[self.idx[prim, k]] + offset*self.vertex_normals[self.idx[prim, k]] for k in range(self.dim)]
        if self.dim == 3:  # samplePointOnTriangle (geometry_helpers.h:182-196)
            u1 = np.sqrt(u[:, 0:1])
            a, b = 1.0 - u1, u[:, 1:2]*u1
            pt = v[0]*a + v[1]*b + v[2]*(1.0 - a - b)
        else:              # samplePointOnLineSegment (geometry_helpers.h:65-75)
            pt = v[0] + u[:, 0:1]*(v[1] - v[0])
        nrm = self._prim_normals(v)
        nrm = nrm/np.maximum(np.linalg.norm(nrm, axis=1, keepdims=True), 1e-30)
        return Samples(pt, nrm, sample_type, weights[prim]/mass, normal_aligned)


class UniformDomainSampler:
    """UniformDomainSampler (domain_samplers.h:22-108): stratified points of the bounding box that fall inside the region.
    inside_solve_region maps an [n, dim] array to a bool array; the reference's demo passes GeometricQueries::insideDomain
    (demo/model_problem.h), whose batched device form is native.Scene.inside_domain (zb_inside_domain)."""

    def __init__(self, inside_solve_region, region_min, region_max, region_volume, seed=0xD0D0):
        self.inside = inside_solve_region
        self.lo, self.hi = np.asarray(region_min, F), np.asarray(region_max, F)
        self.volume = float(region_volume)
        self.rng = np.random.RandomState(seed)

    def generate_samples(self, n):
        dim = len(self.lo)
        extent = self.hi - self.lo
        pdf = 1.0/self.volume
        n_strat = n
        if self.volume > 0:
            n_strat = int(min(n*float(np.prod(extent))*pdf, 1e8))
        s = generate_stratified_samples(n_strat, dim, self.rng)
        pt = (self.lo + extent*s).astype(F)
        pt = pt[self.inside(pt)]
        return Samples(pt, np.zeros_like(pt), abi.IN_DOMAIN, pdf, False)


class NativeBackend:
    """the device: native.Scene for distances and walks, native.Context for the splat"""

    def __init__(self, ctx, scene):
        self.ctx, self.scene, self.dim = ctx, scene, scene.dim

    def distances(self, x, has_absorbing=True, has_reflecting=True):
        da = self.scene.closest_point(abi.ABSORBING, x, want_normal=False)[1] if has_absorbing and len(x) else np.full(len(x), abi.FLT_MAX, F)
        dr = self.scene.closest_point(abi.REFLECTING, x, want_normal=False)[1] if has_reflecting and len(x) else np.full(len(x), abi.FLT_MAX, F)
        return da, dr

    def solve(self, pde, settings, pts, stats):
        return self.scene.solve(abi.WALK_ON_STARS, pde, settings, pts, stats)

    def splat(self, absorption_coeff, samples, eval_pts, **kw):
        return self.ctx.bvc_splat(self.dim, absorption_coeff, samples, eval_pts, **kw)

    def inside_domain(self, x):
        """GeometricQueries::insideDomain on the device: the inside test to hand to UniformDomainSampler"""
        return self.scene.inside_domain(x) if len(x) else np.zeros(0, bool)


class BoundaryValueCachingSolver:
    """bvc::BoundaryValueCachingSolver (boundary_value_caching.h:195-241, 814-925)"""

    def __init__(self, backend, absorbing_sampler, reflecting_sampler, domain_sampler, has_absorbing=True, has_reflecting=True,
                 seed=0xCAC4E):
        self.backend = backend
        self.dim = backend.dim
        self.absorbing_sampler, self.reflecting_sampler, self.domain_sampler = absorbing_sampler, reflecting_sampler, domain_sampler
        self.has_absorbing, self.has_reflecting = has_absorbing, has_reflecting
        self.seed = seed
        self.caches = {}

    def _with_distances(self, s):
        if s is not None and len(s):
            s.dist_a, s.dist_r = self.backend.distances(s.pt, self.has_absorbing, self.has_reflecting)
        return s

    def generate_samples(self, absorbing_cache_size, reflecting_cache_size, domain_cache_size,
                         normal_offset_absorbing, normal_offset_reflecting, solve_double_sided=False):
        """:828-854"""
        c = {}
        a, r = self.absorbing_sampler, self.reflecting_sampler
        if a is not None:
            c["absorbing"] = a.generate_samples(a.get_sample_count(absorbing_cache_size, False), abi.ON_ABSORBING_BOUNDARY, normal_offset_absorbing)
            if solve_double_sided:
                c["absorbing_aligned"] = a.generate_samples(a.get_sample_count(absorbing_cache_size, True), abi.ON_ABSORBING_BOUNDARY,
                                                            normal_offset_absorbing, True)
        if r is not None:
            c["reflecting"] = r.generate_samples(r.get_sample_count(reflecting_cache_size, False), abi.ON_REFLECTING_BOUNDARY, normal_offset_reflecting)
            if solve_double_sided:
                c["reflecting_aligned"] = r.generate_samples(r.get_sample_count(reflecting_cache_size, True), abi.ON_REFLECTING_BOUNDARY,
                                                             normal_offset_reflecting, True)
        if self.domain_sampler is not None and domain_cache_size > 0:
            c["domain"] = self.domain_sampler.generate_samples(domain_cache_size)
        self.caches = {k: self._with_distances(v) for k, v in c.items()}

    def _compute_boundary_estimates(self, s, pde, settings, n_walks_solution, n_walks_gradient, robin_cutoff, pde_eval, stream_base):
        """computeBoundaryEstimates = setEstimationData + WalkOnStars::solve + setEstimatedBoundaryData (:319-343, 567-679),
        unbiased (non finite-difference) variant"""
        if s is None or not len(s):
            return 0
        n = len(s)
        sign = np.where((s.aligned != 0) & bool(settings.solveDoubleSided), F(-1), F(1))[:, None]
        direction = (s.normal*sign).astype(F)
        on_refl = s.type == abi.ON_REFLECTING_BOUNDARY
        if not pde.areRobinConditionsPureNeumann:
            s.robin_coeff = np.where(on_refl, pde_eval.robin_coeff(s.pt, s.normal), 0.0).astype(F)
        wants_gradient = (~on_refl) | (np.abs(s.robin_coeff) > robin_cutoff)
        pts = abi.HostSamplePoints(s.pt, s.dist_a, s.dist_r, nWalks=np.where(wants_gradient, n_walks_gradient, n_walks_solution),
                                   estimationQuantity=np.where(wants_gradient, abi.SOLUTION_AND_GRADIENT, abi.SOLUTION),
                                   type=s.type, normal=s.normal, seed=self.seed, streams=np.arange(n, dtype=np.uint64) + np.uint64(stream_base))
        pts.directionForDerivative = np.ascontiguousarray(direction)
        pts.estimateBoundaryNormalAligned = np.ascontiguousarray(s.aligned)
        st = abi.HostStatistics(n, self.dim)
        steps = self.backend.solve(pde, settings, pts, st)
        s.solution = st.solutionMean.copy()
        derivative = st.totalDerivativeContribution/np.maximum(1, st.nSolutionEstimates)   # getEstimatedDerivative (common.h:341-345)
        if not settings.ignoreReflectingBoundaryContribution:
            h = pde_eval.robin(s.pt, s.normal)
            if pde.areRobinConditionsPureNeumann:
                s.normal_derivative = np.where(on_refl, h, s.normal_derivative).astype(F)
            else:
                s.contribution = np.where(on_refl, h, s.contribution).astype(F)
                s.normal_derivative = np.where(on_refl & (np.abs(s.robin_coeff) > robin_cutoff), derivative, s.normal_derivative).astype(F)
        s.normal_derivative = np.where(~on_refl, derivative, s.normal_derivative).astype(F)
        return steps

    def compute_sample_estimates(self, pde, settings, n_walks_solution, n_walks_gradient, robin_cutoff, pde_eval):
        """:857-883. pde_eval gives the host values of the PDE callbacks at arrays of points (abi.PdeDesc is data, so
        they are closed forms): .robin(x, n), .robin_coeff(x, n), .source(x). Returns the walk steps taken."""
        steps = 0
        for j, key in enumerate(("absorbing", "absorbing_aligned", "reflecting", "reflecting_aligned")):
            steps += self._compute_boundary_estimates(self.caches.get(key), pde, settings, n_walks_solution, n_walks_gradient,
                                                      robin_cutoff, pde_eval, stream_base=(j + 1) << 32)
        d = self.caches.get("domain")
        if d is not None and len(d):   # setSourceValues (:345-367)
            d.contribution = np.asarray(pde_eval.source(d.pt), F)
        return steps

    def cache(self):
        """all cache samples as one abi.HostBvcSamples"""
        s = Samples.concatenate([self.caches.get(k) for k in ("absorbing", "absorbing_aligned", "reflecting", "reflecting_aligned", "domain")])
        if s is None:
            z = np.zeros(0, F)
            return abi.HostBvcSamples(np.zeros((0, self.dim), F), np.zeros((0, self.dim), F), np.zeros(0, np.int32), z, z, z, z, z)
        return abi.HostBvcSamples(s.pt, s.normal, s.type, s.pdf, s.solution, s.normal_derivative, s.contribution, s.robin_coeff, s.aligned)

    def splat(self, pde, eval_pts, radius_clamp, kernel_regularization, robin_cutoff, cutoff_a, cutoff_r):
        """:886-910; eval_pts is an abi.HostBvcEvalPoints, updated in place"""
        return self.backend.splat(float(pde.absorptionCoeff), self.cache(), eval_pts, radius_clamp=radius_clamp,
                                  kernel_regularization=kernel_regularization, robin_cutoff=robin_cutoff, cutoff_a=cutoff_a, cutoff_r=cutoff_r)

    def estimate_solution_near_boundary(self, pde, settings, eval_pts, cutoff_a, cutoff_r, n_walks, normals=None):
        """:913-925 -> :505-562: evaluation points closer to a boundary than its cutoff get a plain WoSt estimate, recorded
        as the single estimate of the statistics bucket their type selects"""
        near = (eval_pts.distToAbsorbingBoundary < cutoff_a) | (eval_pts.distToReflectingBoundary < cutoff_r)
        idx = np.where(near)[0]
        if not len(idx):
            return 0
        nrm = np.zeros((len(idx), self.dim), F) if normals is None else np.asarray(normals, F)[idx]
        pts = abi.HostSamplePoints(eval_pts.pt[idx], eval_pts.distToAbsorbingBoundary[idx], eval_pts.distToReflectingBoundary[idx],
                                   nWalks=n_walks, estimationQuantity=abi.SOLUTION, type=eval_pts.type[idx], normal=nrm, seed=self.seed,
                                   streams=idx.astype(np.uint64) + np.uint64(7 << 32))
        st = abi.HostStatistics(len(idx), self.dim)
        steps = self.backend.solve(pde, settings, pts, st)
        # a point inside both cutoffs is estimated twice by the reference (:913-925), once per call; one estimate is kept here
        bucket = np.where(eval_pts.type[idx] == abi.ON_ABSORBING_BOUNDARY, 0, np.where(eval_pts.type[idx] == abi.ON_REFLECTING_BOUNDARY, 2, 4))
        eval_pts.bucketCount[idx, bucket] += 1
        eval_pts.bucketSum[idx, bucket, 0] += st.solutionMean
        mean = eval_pts.bucket_mean()
        eval_pts.solution[:] = mean[:, :, 0].sum(axis=1)
        eval_pts.gradient[:] = mean[:, :, 1:].sum(axis=1)
        return steps


class ManufacturedPdeEval:
    """host values of the data-described PDE's callbacks (abi.PdeDesc.manufactured; include/zombie_b200/analytic.h) for
    the families the workloads use, vectorised: what pde.robin / robinCoeff / source return in the reference"""

    def __init__(self, u, grad_u, lap_u, sigma, mu):
        self.u, self.grad_u, self.lap_u, self.sigma, self.mu = u, grad_u, lap_u, float(sigma), float(mu)

    def source(self, x):
        return (self.sigma*self.u(x) - self.lap_u(x)).astype(F)

    def robin(self, x, n):
        return ((self.grad_u(x)*n).sum(axis=1) + self.mu*self.u(x)).astype(F)

    def robin_coeff(self, x, n):
        return np.full(len(x), self.mu, F)
