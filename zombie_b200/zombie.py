"""The reference's Python surface (python/zombie/__init__.py:128-330 over python/zombie_bindings.cpp) for the walk path,
on top of the C ABI: `from zombie_b200 import zombie` and the calls of the reference's demo (demo/demo.py:124-540) —
zombie.Utils.load_boundary_mesh / normalize / flip_orientation / compute_bounding_box / partition_boundary_mesh,
zombie.Core.PDE, zombie.Utils.get_dense_grid_*_callback, zombie.Core.GeometricQueries, the Fcpw*BoundaryHandler classes and
populate_geometric_queries_for_*_boundary, zombie.Utils.compute_dist_to_boundary, zombie.Solvers.WalkSettings / SamplePoint /
SamplePointList / create_sample_statistics_list / WalkOnSpheres / WalkOnStars — keep their names, argument order and
`dim=` / `channels=` keywords. What differs, because walks run on the device:

* PDE callbacks are DATA: the objects returned by get_dense_grid_*_callback and get_constant_*_callback (the only kinds
  the demo uses) describe a zb_grid_view or a constant; an arbitrary Python callable is accepted only for
  has_reflecting_boundary_conditions, which is evaluated on the host (partition_boundary_mesh). Assigning any other callable
  to source / dirichlet / robin / robin_coeff raises at solve time.
* channels > 1 (the reference's T = Array<float, channels>) runs as ONE set of walks with one data descriptor and one set of
  statistics per channel (zb_solve_channels): constant callbacks take `channels` values, dense grids `channels` values per
  cell (last axis), and the statistics getters return one value per channel.
* walk_state_callback / terminal_contribution_callback / report_progress are accepted and ignored (no host callbacks inside
  walks); run_single_threaded is ignored.
The list types are Python lists of numpy rows; convert_list_to_numpy_array gives the array."""
import math

import numpy as np

from . import abi, native

_VALID_DIMS = (2, 3)


def _check_dim(dim):
    if dim not in _VALID_DIMS:
        raise ValueError("dim must be one of [2, 3], got %r" % (dim,))


def _check_channels(channels):
    if not isinstance(channels, (int, np.integer)) or not 1 <= channels <= 256:
        raise ValueError("channels must be an integer in 1..256, got %r" % (channels,))


# ---- list types -------------------------------------------------------------------------------
class FloatList(list):
    pass


class IntList(list):
    pass


class UintList(list):
    pass


class BoolList(list):
    pass


class Float2List(list):
    pass


class Float3List(list):
    pass


class Int2List(list):
    pass


class Int3List(list):
    pass


def FloatNList(*args, dim):
    _check_dim(dim)
    return (Float2List if dim == 2 else Float3List)(*args)


def IntNList(*args, dim):
    _check_dim(dim)
    return (Int2List if dim == 2 else Int3List)(*args)


def convert_list_to_numpy_array(values):
    return np.asarray(values)


def _points(values, dim, dtype=np.float32):
    a = np.asarray(values, dtype=dtype)
    return np.ascontiguousarray(a.reshape(-1, dim))


def _refill(dst, rows):
    del dst[:]
    dst.extend(rows)


_default_context = None


def _context():
    global _default_context
    if _default_context is None:
        _default_context = native.Context(0)
    return _default_context


# ---- PDE data ---------------------------------------------------------------------------------
class ConstantCallback:
    """what Core.get_constant_*_callback returns"""

    def __init__(self, value, channels=1):
        v = np.asarray(value, np.float64).reshape(-1)
        if v.size not in (1, channels):
            raise ValueError("constant callback: %d values for %d channels" % (v.size, channels))
        self.values = np.broadcast_to(v, (channels,)).copy()
        self.value = float(self.values[0])

    def channel(self, c):
        return ConstantCallback(self.values[c] if len(self.values) > 1 else self.values[0])

    def __call__(self, *args):
        return self.value if len(self.values) == 1 else self.values.copy()


class DenseGridCallback:
    """what Utils.get_dense_grid_*_callback returns: DenseGrid<float, 1, DIM> (zombie/utils/dense_grid.h:29-72) — cell-centred
    samples over [grid_min, grid_max], last axis fastest, nearest-cell lookup or multilinear interpolation with edge clamping —
    as an abi.HostGrid: the same samples on the lattice of the cell centres."""

    def __init__(self, data, shape, grid_min, grid_max, dim, enable_interpolation=False, aligned_data=None, channels=1):
        shape = [int(s) for s in np.asarray(shape).reshape(-1)[:dim]]
        self.dim = dim
        # DenseGrid<float, CHANNELS, DIM>: `channels` values per cell, the channel index fastest
        def split(d):
            v = np.asarray(d, np.float32).reshape(-1)
            if v.size != int(np.prod(shape))*channels:
                raise ValueError("dense grid: %d values for shape %s x %d channels" % (v.size, shape, channels))
            v = v.reshape(-1, channels)
            return [self._grid(np.ascontiguousarray(v[:, c]), shape, grid_min, grid_max, dim, enable_interpolation) for c in range(channels)]
        self.grids = split(data)
        self.aligned_grids = None if aligned_data is None else split(aligned_data)
        self.grid = self.grids[0]
        self.aligned_grid = None if self.aligned_grids is None else self.aligned_grids[0]

    def channel(self, c):
        o = object.__new__(DenseGridCallback)
        o.dim = self.dim
        o.grids = [self.grids[c]]
        o.aligned_grids = None if self.aligned_grids is None else [self.aligned_grids[c]]
        o.grid, o.aligned_grid = o.grids[0], (None if o.aligned_grids is None else o.aligned_grids[0])
        return o

    @staticmethod
    def _grid(data, shape, grid_min, grid_max, dim, enable_interpolation):
        v = np.asarray(data, np.float32).reshape(-1)
        if v.size != int(np.prod(shape)):
            raise ValueError("dense grid: %d values for shape %s" % (v.size, shape))
        v = v.reshape(shape)                                   # [nx, ny(, nz)], last axis fastest
        v = v.transpose(*reversed(range(dim)))                  # [(nz,) ny, nx]: x fastest (zb_grid_view)
        lo_b, hi_b = np.asarray(grid_min, np.float64)[:dim], np.asarray(grid_max, np.float64)[:dim]
        spacing = (hi_b - lo_b)/np.asarray(shape, np.float64)
        lo, hi = lo_b + 0.5*spacing, hi_b - 0.5*spacing        # the cell centres
        return abi.HostGrid(v, lo, hi, abi.GRID_LINEAR if enable_interpolation else abi.GRID_NEAREST)

    def __call__(self, x, *args):
        """host evaluation (nearest cell or multilinear), for partition_boundary_mesh and for tests"""
        g = self.grid
        if len(args) >= 1 and isinstance(args[-1], (bool, np.bool_)) and args[-1] and self.aligned_grid is not None:
            g = self.aligned_grid
        x = np.asarray(x, np.float64).reshape(-1)[:self.dim]
        nz, ny, nx = g.values.shape
        n = (nx, ny, nz)[:self.dim]
        t = [(x[k] - g.lo[k])*((n[k] - 1)/(g.hi[k] - g.lo[k]) if n[k] > 1 else 0.0) for k in range(self.dim)]
        if g.interp == abi.GRID_NEAREST:
            idx = [int(min(max(math.floor(t[k] + 0.5), 0), n[k] - 1)) for k in range(self.dim)] + [0]*(3 - self.dim)
            return float(g.values[idx[2], idx[1], idx[0]])
        acc = 0.0
        i0, w = [], []
        for k in range(self.dim):
            c = min(max(t[k], 0.0), n[k] - 1.0)
            i = min(int(math.floor(c)), max(n[k] - 2, 0))
            i0.append(i)
            w.append(c - i if n[k] > 1 else 0.0)
        for corner in range(1 << self.dim):
            idx, wt = [0, 0, 0], 1.0
            for k in range(self.dim):
                bit = (corner >> k) & 1
                idx[k] = min(i0[k] + bit, n[k] - 1)
                wt *= w[k] if bit else 1.0 - w[k]
            acc += wt*float(g.values[idx[2], idx[1], idx[0]])
        return acc


class _PDE:
    """zombie::PDE<float, DIM> (core/pde.h:20-45) with the attribute names of the binding (zombie_bindings.h)"""

    def __init__(self, dim, channels=1):
        self.dim = dim
        self.channels = channels
        self.absorption_coeff = 0.0
        self.is_source_constant = False
        self.are_robin_conditions_pure_neumann = True
        self.are_robin_coeffs_nonnegative = True
        self.source = ConstantCallback(0.0)
        self.dirichlet = ConstantCallback(0.0)
        self.robin = ConstantCallback(0.0)
        self.robin_coeff = ConstantCallback(0.0)
        self.has_reflecting_boundary_conditions = lambda x: False
        self.kelvin_origin = None   # set by KelvinTransform-style callers: solve on the inverted domain

    def descriptors(self):
        """one zb_pde_desc per value channel (zb_solve_channels): what steers the walk is shared, the data are the channel's"""
        out = []
        for c in range(self.channels):
            one = _PDE(self.dim)
            for f in ("absorption_coeff", "is_source_constant", "are_robin_conditions_pure_neumann", "are_robin_coeffs_nonnegative",
                      "kelvin_origin", "robin_coeff"):
                setattr(one, f, getattr(self, f))
            for f in ("source", "dirichlet", "robin"):
                cb = getattr(self, f)
                setattr(one, f, cb.channel(c) if isinstance(cb, (ConstantCallback, DenseGridCallback)) and self.channels > 1 else cb)
            out.append(one.descriptor())
        return out

    def descriptor(self):
        """the zb_pde_desc of this PDE"""
        d = abi.PdeDesc()
        d.absorptionCoeff = float(self.absorption_coeff)
        d.isSourceConstant = int(bool(self.is_source_constant))
        d.areRobinConditionsPureNeumann = int(bool(self.are_robin_conditions_pure_neumann))
        d.areRobinCoeffsNonnegative = int(bool(self.are_robin_coeffs_nonnegative))
        d.mode = abi.PDE_CONSTANT
        grids = {}
        for name, field, key in (("source", "sourceValue", "source"), ("dirichlet", "dirichletValue", "dirichlet"),
                                 ("robin", "robinValue", "robin"), ("robin_coeff", "robinCoeff", "robinCoeff")):
            cb = getattr(self, name)
            if isinstance(cb, ConstantCallback):
                setattr(d, field, cb.value)
            elif isinstance(cb, DenseGridCallback):
                grids[key] = cb.grid
                if name == "robin" and cb.aligned_grid is not None:
                    grids["robinAligned"] = cb.aligned_grid
                if name == "dirichlet" and cb.aligned_grid is not None:
                    grids["dirichletAligned"] = cb.aligned_grid
            else:
                raise TypeError("PDE.%s must come from get_dense_grid_*_callback or get_constant_*_callback: walks run on the "
                                "device and cannot call back into Python" % name)
        if grids:
            d.set_grids(**grids)
        if self.kelvin_origin is not None:
            d.kelvin = 1
            for k in range(self.dim):
                d.kelvinOrigin[k] = float(self.kelvin_origin[k])
        return d


class _GeometricQueries:
    """zombie::GeometricQueries<DIM> (core/geometric_queries.h:59-147) bound to a device scene that is built when first needed"""

    def __init__(self, domain_is_watertight, domain_min, domain_max, dim):
        self.dim = dim
        self.domain_is_watertight = bool(domain_is_watertight)
        self.domain_min = np.asarray(domain_min, np.float32).reshape(-1)[:dim].copy()
        self.domain_max = np.asarray(domain_max, np.float32).reshape(-1)[:dim].copy()
        self.has_non_empty_absorbing_boundary = False
        self.has_non_empty_reflecting_boundary = False
        self._absorbing = None
        self._reflecting = None
        self._scene = None

    def _invalidate(self):
        if self._scene is not None:
            self._scene.close()
        self._scene = None

    def scene(self):
        if self._scene is None:
            kw = dict(watertight=self.domain_is_watertight, domain_min=self.domain_min, domain_max=self.domain_max)
            if self._absorbing is not None:
                kw["absorbing"] = (self._absorbing.positions, self._absorbing.indices)
            r = self._reflecting
            if r is not None:
                kw.update(reflecting=(r.positions, r.indices), reflecting_kind=r.kind, solve_double_sided=r.solve_double_sided,
                          silhouette_precision=r.silhouette_precision)
                if r.kind == abi.REFLECTING_ROBIN:
                    kw.update(min_robin=r.min_robin, max_robin=r.max_robin)
            self._scene = native.Scene(_context(), self.dim, **kw)
        return self._scene

    # the single-point callbacks of the reference, each a batched query of size one
    def compute_dist_to_absorbing_boundary(self, x, compute_signed_distance=False):
        if not self.has_non_empty_absorbing_boundary:
            return float(abi.FLT_MAX)
        return float(self.scene().closest_point(abi.ABSORBING, _points(x, self.dim), want_normal=False)[1][0])

    def compute_dist_to_reflecting_boundary(self, x, compute_signed_distance=False):
        if not self.has_non_empty_reflecting_boundary:
            return float(abi.FLT_MAX)
        return float(self.scene().closest_point(abi.REFLECTING, _points(x, self.dim), want_normal=False)[1][0])

    def compute_dist_to_boundary(self, x, compute_signed_distance=False):
        return min(self.compute_dist_to_absorbing_boundary(x), self.compute_dist_to_reflecting_boundary(x))

    def inside_domain(self, x):
        if not self.domain_is_watertight:
            return True
        if not (self.has_non_empty_absorbing_boundary or self.has_non_empty_reflecting_boundary):
            return False
        return bool(self.scene().inside_domain(_points(x, self.dim))[0])

    def inside_domain_batch(self, x):
        x = _points(x, self.dim)
        if not self.domain_is_watertight:
            return np.ones(len(x), bool)
        return self.scene().inside_domain(x)

    def inside_bounding_domain(self, x):
        x = np.asarray(x, np.float32).reshape(-1)[:self.dim]
        return bool(((x >= self.domain_min) & (x <= self.domain_max)).all())

    def outside_bounding_domain(self, x):
        return not self.inside_bounding_domain(x)


class _BoundaryHandler:
    def __init__(self, dim, kind):
        self.dim, self.kind = dim, kind
        self.positions = np.zeros((0, dim), np.float32)
        self.indices = np.zeros((0, dim), np.int32)
        self.solve_double_sided = False
        self.silhouette_precision = 1e-3
        self.min_robin = self.max_robin = None
        self._queries = []

    def _set_mesh(self, positions, indices):
        self.positions = _points(positions, self.dim)
        self.indices = _points(indices, self.dim, np.int32)


class _DirichletHandler(_BoundaryHandler):
    """FcpwDirichletBoundaryHandler (fcpw_geometric_queries.h:75-88)"""

    def __init__(self, dim):
        super().__init__(dim, abi.REFLECTING_NONE)

    def build_acceleration_structure(self, positions, indices, build_bvh=True, enable_bvh_vectorization=False, enable_gpu=False):
        self._set_mesh(positions, indices)


class IgnoreCandidateSilhouette:
    """getIgnoreCandidateSilhouetteCallback (fcpw_geometric_queries.h:1029-1038) as data"""

    def __init__(self, solve_double_sided=False, silhouette_precision=1e-3):
        self.solve_double_sided, self.silhouette_precision = bool(solve_double_sided), float(silhouette_precision)


class BranchTraversalWeight:
    """getBranchTraversalWeightCallback (fcpw_geometric_queries.h:1391-1398); the device traversal builds in min_radial_dist = 1e-2"""

    def __init__(self, min_radial_dist=1e-2):
        if min_radial_dist != 1e-2:
            raise ValueError("branch traversal weight: only min_radial_dist = 1e-2 is built into the device traversal")
        self.min_radial_dist = min_radial_dist


class _NeumannHandler(_BoundaryHandler):
    """FcpwNeumannBoundaryHandler (fcpw_geometric_queries.h:90-105)"""

    def __init__(self, dim):
        super().__init__(dim, abi.REFLECTING_NEUMANN)

    def build_acceleration_structure(self, positions, indices, ignore_candidate_silhouette=None, build_bvh=True,
                                     enable_bvh_vectorization=False, enable_gpu=False):
        self._set_mesh(positions, indices)
        ics = ignore_candidate_silhouette or IgnoreCandidateSilhouette()
        self.solve_double_sided, self.silhouette_precision = ics.solve_double_sided, ics.silhouette_precision


class _RobinHandler(_BoundaryHandler):
    """FcpwRobinBoundaryHandler (fcpw_geometric_queries.h:107-133)"""

    def __init__(self, dim):
        super().__init__(dim, abi.REFLECTING_ROBIN)

    def build_acceleration_structure(self, positions, indices, ignore_candidate_silhouette, min_robin_coeff_values,
                                     max_robin_coeff_values, build_bvh=True, enable_bvh_vectorization=False, enable_gpu=False):
        self._set_mesh(positions, indices)
        ics = ignore_candidate_silhouette or IgnoreCandidateSilhouette()
        self.solve_double_sided, self.silhouette_precision = ics.solve_double_sided, ics.silhouette_precision
        self.min_robin = np.ascontiguousarray(min_robin_coeff_values, np.float32)
        self.max_robin = np.ascontiguousarray(max_robin_coeff_values, np.float32)
        if len(self.min_robin) != len(self.indices) or len(self.max_robin) != len(self.indices):
            raise ValueError("one minimum and one maximum Robin coefficient per reflecting primitive")

    def update_coefficient_values(self, min_robin_coeff_values, max_robin_coeff_values):
        """FcpwRobinBoundaryHandler::updateCoefficientValues: refitted on the device for every scene already built"""
        self.min_robin = np.ascontiguousarray(min_robin_coeff_values, np.float32)
        self.max_robin = np.ascontiguousarray(max_robin_coeff_values, np.float32)
        for q in self._queries:
            if q._scene is not None:
                q._scene.update_robin_coefficients(self.min_robin, self.max_robin)


# ---- solver-side types --------------------------------------------------------------------------
class SampleType:
    InDomain, OnAbsorbingBoundary, OnReflectingBoundary = abi.IN_DOMAIN, abi.ON_ABSORBING_BOUNDARY, abi.ON_REFLECTING_BOUNDARY


class EstimationQuantity:
    Solution, SolutionAndGradient, Skip = abi.SOLUTION, abi.SOLUTION_AND_GRADIENT, abi.SKIP


def WalkSettings(epsilon_shell_for_absorbing_boundary, epsilon_shell_for_reflecting_boundary, *args):
    """zombie::WalkSettings, both constructors (point_estimation/common.h:21-73): (epsA, epsR, maxWalkLength, solveDoubleSided)
    or the sixteen-argument form the demo uses"""
    if len(args) == 2:
        return abi.WalkSettings.default(epsilon_shell_for_absorbing_boundary, epsilon_shell_for_reflecting_boundary, int(args[0]), bool(args[1]))
    if len(args) != 14:
        raise TypeError("WalkSettings takes 4 or 16 arguments")
    (silhouette_precision, russian_roulette_threshold, splitting_threshold, max_walk_length, steps_before_applying_tikhonov,
     steps_before_using_maximal_spheres, solve_double_sided, use_gradient_control_variates, use_gradient_antithetic_variates,
     use_cosine_sampling_for_derivatives, ignore_absorbing, ignore_reflecting, ignore_source, print_logs) = args
    splitting = float(abi.FLT_MAX) if not np.isfinite(splitting_threshold) else float(splitting_threshold)
    return abi.WalkSettings(epsilon_shell_for_absorbing_boundary, epsilon_shell_for_reflecting_boundary, silhouette_precision,
                            russian_roulette_threshold, min(splitting, float(abi.FLT_MAX)), int(max_walk_length),
                            int(steps_before_applying_tikhonov), int(steps_before_using_maximal_spheres), int(bool(solve_double_sided)),
                            int(bool(use_gradient_control_variates)), int(bool(use_gradient_antithetic_variates)),
                            int(bool(use_cosine_sampling_for_derivatives)), int(bool(ignore_absorbing)), int(bool(ignore_reflecting)),
                            int(bool(ignore_source)), int(bool(print_logs)))


class _SamplePoint:
    """zombie::SamplePoint<float, DIM> (point_estimation/common.h:152-195)"""
    _next_stream = 0

    def __init__(self, pt, normal, sample_type, estimation_quantity, pdf, dist_to_absorbing_boundary, dist_to_reflecting_boundary, dim):
        self.pt = np.asarray(pt, np.float32).reshape(-1)[:dim].copy()
        self.normal = np.asarray(normal, np.float32).reshape(-1)[:dim].copy()
        self.type = int(sample_type)
        self.estimation_quantity = int(estimation_quantity)
        self.pdf = float(pdf)
        self.dist_to_absorbing_boundary = float(dist_to_absorbing_boundary)
        self.dist_to_reflecting_boundary = float(dist_to_reflecting_boundary)
        self.direction_for_derivative = np.eye(dim, dtype=np.float32)[0]
        self.first_sphere_radius = 0.0
        self.estimate_boundary_normal_aligned = False
        # the reference seeds every sample point's pcg32 from the clock (sampling.h:36-41); here: a stream per point
        self.rng_state, self.rng_inc = None, None
        self.stream = _SamplePoint._next_stream
        _SamplePoint._next_stream += 1


class _SampleStatistics:
    """zombie::SampleStatistics<float, DIM> (point_estimation/common.h:201-395): a view of row i of the SoA statistics"""

    def __init__(self, owner, i, others=()):
        self._o, self._i = owner, i
        self._all = (owner,) + tuple(others)   # one HostStatistics per value channel

    def _per_channel(self, f):
        """a float for scalar problems, one entry per channel (the reference's T = Array<float, channels>) otherwise"""
        if len(self._all) == 1:
            return f(self._o)
        return np.array([f(o) for o in self._all])

    def get_estimated_solution(self):
        return self._per_channel(lambda o: float(o.solutionMean[self._i]))

    def get_estimated_solution_variance(self):
        n = int(self._o.nSolutionEstimates[self._i])
        return self._per_channel(lambda o: float(o.solutionM2[self._i]/max(1, n - 1)))

    def get_estimated_gradient(self):
        """DIM entries (of T: with several channels, rows = coordinates, columns = channels)"""
        if len(self._all) == 1:
            return self._o.gradientMean[self._i].copy()
        return np.stack([o.gradientMean[self._i] for o in self._all], axis=1)

    def get_estimated_gradient_variance(self):
        n = int(self._o.nGradientEstimates[self._i])
        if len(self._all) == 1:
            return self._o.gradientM2[self._i]/max(1, n - 1)
        return np.stack([o.gradientM2[self._i]/max(1, n - 1) for o in self._all], axis=1)

    def get_estimated_derivative(self):
        n = max(1, int(self._o.nSolutionEstimates[self._i]))
        return self._per_channel(lambda o: float(o.totalDerivativeContribution[self._i]/n))

    def get_solution_estimate_count(self):
        return int(self._o.nSolutionEstimates[self._i])

    def get_gradient_estimate_count(self):
        return int(self._o.nGradientEstimates[self._i])

    def get_mean_walk_length(self):
        return float(self._o.totalWalkLength[self._i]/max(1, int(self._o.nSolutionEstimates[self._i])))


class _SampleStatisticsList:
    def __init__(self, n, dim, channels=1):
        self.hosts = [abi.HostStatistics(n, dim) for _ in range(channels)]
        self.host = self.hosts[0]

    def __len__(self):
        return self.host.n

    def __getitem__(self, i):
        if not -self.host.n <= i < self.host.n:
            raise IndexError(i)
        return _SampleStatistics(self.host, i % self.host.n, self.hosts[1:])


class _WalkSolver:
    def __init__(self, geometric_queries, estimator, walk_state_callback=None, terminal_contribution_callback=None):
        self.queries, self.estimator = geometric_queries, estimator
        self.last_walk_steps = 0

    def solve(self, pde, walk_settings, n_walks, sample_pts, sample_statistics, run_single_threaded=False, report_progress=None):
        """WalkOn{Spheres,Stars}::solve, the vector overload (walk_on_spheres.h:132-168, walk_on_stars.h:151-187): one zb_solve"""
        q = self.queries
        n = len(sample_pts)
        if len(n_walks) != n or len(sample_statistics) != n:
            raise ValueError("n_walks, sample_pts and sample_statistics must have the same length")
        if n == 0:
            return
        dim = q.dim
        seeds = np.array([p.stream for p in sample_pts], np.uint64)
        pts = abi.HostSamplePoints(np.stack([p.pt for p in sample_pts]),
                                   np.array([p.dist_to_absorbing_boundary for p in sample_pts], np.float32),
                                   np.array([p.dist_to_reflecting_boundary for p in sample_pts], np.float32),
                                   nWalks=np.asarray(n_walks, np.int32),
                                   estimationQuantity=np.array([p.estimation_quantity for p in sample_pts], np.int32),
                                   type=np.array([p.type for p in sample_pts], np.int32),
                                   normal=np.stack([p.normal for p in sample_pts]), streams=seeds)
        pts.directionForDerivative = np.ascontiguousarray(np.stack([p.direction_for_derivative for p in sample_pts]), np.float32)
        pts.estimateBoundaryNormalAligned = np.array([p.estimate_boundary_normal_aligned for p in sample_pts], np.uint8)
        pts.firstSphereRadius = np.array([p.first_sphere_radius for p in sample_pts], np.float32)
        for i, p in enumerate(sample_pts):   # generators that already advanced (progressive solves) continue where they stopped
            if p.rng_state is not None:
                pts.rngState[i], pts.rngInc[i] = p.rng_state, p.rng_inc
        if len(sample_statistics.hosts) != pde.channels:
            raise ValueError("the PDE has %d value channels, the statistics %d" % (pde.channels, len(sample_statistics.hosts)))
        if pde.channels == 1:
            self.last_walk_steps = q.scene().solve(self.estimator, pde.descriptor(), walk_settings, pts, sample_statistics.host)
        else:   # T = Array<float, channels>: one set of walks for all channels
            self.last_walk_steps = q.scene().solve_channels(self.estimator, pde.descriptors(), walk_settings, pts, sample_statistics.hosts)
        for i, p in enumerate(sample_pts):
            p.rng_state, p.rng_inc = int(pts.rngState[i]), int(pts.rngInc[i])
            p.first_sphere_radius = float(pts.firstSphereRadius[i])
        if report_progress is not None:
            try:
                report_progress(n, 0)
            except TypeError:
                pass


# ---- the four namespaces of the reference's module ---------------------------------------------------
class Core:
    @staticmethod
    def PDE(*, dim, channels=1):
        _check_dim(dim); _check_channels(channels)
        return _PDE(dim, channels)

    @staticmethod
    def GeometricQueries(domain_is_watertight=True, domain_min=None, domain_max=None, *, dim):
        _check_dim(dim)
        lo = np.full(dim, -abi.FLT_MAX, np.float32) if domain_min is None else domain_min
        hi = np.full(dim, abi.FLT_MAX, np.float32) if domain_max is None else domain_max
        return _GeometricQueries(domain_is_watertight, lo, hi, dim)

    @staticmethod
    def get_constant_source_callback(value, *, dim, channels=1):
        _check_dim(dim); _check_channels(channels)
        return ConstantCallback(value, channels)

    get_constant_dirichlet_callback = get_constant_source_callback
    get_constant_robin_callback = get_constant_source_callback

    @staticmethod
    def get_constant_robin_coefficient_callback(value, *, dim):
        _check_dim(dim)
        return ConstantCallback(value)

    @staticmethod
    def get_constant_indicator_callback(value, *, dim):
        _check_dim(dim)
        return lambda x, v=bool(value): v


class Solvers:
    SampleType = SampleType
    EstimationQuantity = EstimationQuantity
    WalkSettings = staticmethod(WalkSettings)

    @staticmethod
    def SamplePoint(pt, normal, sample_type, estimation_quantity, pdf, dist_to_absorbing_boundary, dist_to_reflecting_boundary, *, dim, channels=1):
        _check_dim(dim); _check_channels(channels)
        return _SamplePoint(pt, normal, sample_type, estimation_quantity, pdf, dist_to_absorbing_boundary, dist_to_reflecting_boundary, dim)

    @staticmethod
    def SamplePointList(points=(), *, dim, channels=1):
        _check_dim(dim); _check_channels(channels)
        return list(points)

    @staticmethod
    def SampleStatisticsList(n=0, *, dim, channels=1):
        _check_dim(dim); _check_channels(channels)
        return _SampleStatisticsList(n, dim, channels)

    @staticmethod
    def create_sample_statistics_list(n, *, dim, channels=1):
        _check_dim(dim); _check_channels(channels)
        return _SampleStatisticsList(n, dim, channels)

    @staticmethod
    def WalkOnSpheres(geometric_queries, walk_state_callback=None, terminal_contribution_callback=None, *, dim, channels=1):
        _check_dim(dim); _check_channels(channels)
        return _WalkSolver(geometric_queries, abi.WALK_ON_SPHERES)

    @staticmethod
    def WalkOnStars(geometric_queries, walk_state_callback=None, terminal_contribution_callback=None, *, dim, channels=1):
        _check_dim(dim); _check_channels(channels)
        return _WalkSolver(geometric_queries, abi.WALK_ON_STARS)

    @staticmethod
    def get_empty_walk_state_callback(*, dim, channels=1):
        return None

    @staticmethod
    def get_empty_terminal_contribution_callback(*, dim, channels=1):
        return None


class Utils:
    @staticmethod
    def load_boundary_mesh(obj_file, positions, indices, *, dim):
        """loadBoundaryMesh (zombie/utils/fcpw_geometric_queries.h): 'v' vertices and 'l' polylines (2D) or 'f' faces, fan-triangulated (3D)"""
        _check_dim(dim)
        pos, idx = [], []
        with open(obj_file) as f:
            for line in f:
                t = line.split()
                if not t:
                    continue
                if t[0] == "v":
                    pos.append(np.array([float(v) for v in t[1:1 + dim]], np.float32))
                elif t[0] == "f" and dim == 3:
                    v = [int(a.split("/")[0]) - 1 for a in t[1:]]
                    idx += [np.array([v[0], v[i], v[i + 1]], np.int32) for i in range(1, len(v) - 1)]
                elif t[0] == "l" and dim == 2:
                    v = [int(a) - 1 for a in t[1:]]
                    idx += [np.array([v[i], v[i + 1]], np.int32) for i in range(len(v) - 1)]
        _refill(positions, pos)
        _refill(indices, idx)

    @staticmethod
    def normalize(positions, *, dim):
        """normalize (fcpw_geometric_queries.h:255-273): centre of mass to the origin, farthest vertex to distance 1"""
        p = _points(positions, dim)
        p = p - p.mean(axis=0, dtype=np.float32)
        p = p/np.float32(np.linalg.norm(p, axis=1).max())
        _refill(positions, list(p.astype(np.float32)))

    @staticmethod
    def apply_shift(shift, positions, *, dim):
        p = _points(positions, dim) + np.asarray(shift, np.float32)[:dim]
        _refill(positions, list(p))

    @staticmethod
    def flip_orientation(indices, *, dim):
        i = _points(indices, dim, np.int32).copy()
        i[:, [0, 1]] = i[:, [1, 0]]
        _refill(indices, list(i))

    @staticmethod
    def compute_bounding_box(positions, make_square, scale, *, dim):
        """computeBoundingBox (fcpw_geometric_queries.h:292-317)"""
        p = _points(positions, dim)
        cm = (np.cumsum(p, axis=0, dtype=np.float32)[-1]/np.float32(len(p))).astype(np.float32)   # summed in order, like the reference
        q = (p - cm)*np.float32(scale) + cm
        # fcpw::BoundingBox::expandToInclude pads every point by epsilon (fcpw/core/bounding_volumes.h:45-49): geometry lying ON the
        # box (the demo's engine outline does) is inside it for the exact comparisons of insideBoundingDomain
        eps = np.float32(np.finfo(np.float32).eps)
        lo, hi = (q - eps).min(axis=0), (q + eps).max(axis=0)
        if make_square:
            centre, half = 0.5*(lo + hi), np.float32(0.5*(hi - lo).max())
            lo, hi = centre - half, centre + half
        return lo.astype(np.float32), hi.astype(np.float32)

    @staticmethod
    def partition_boundary_mesh(on_reflecting_boundary, positions, indices, absorbing_positions, absorbing_indices,
                                reflecting_positions, reflecting_indices, *, dim):
        """partitionBoundaryMesh (fcpw_geometric_queries.h:460-510): primitives by the predicate at their midpoints, vertices
        re-indexed in order of first use"""
        p, idx = _points(positions, dim), _points(indices, dim, np.int32)
        mid = p[idx].mean(axis=1, dtype=np.float32)
        mask = np.array([bool(on_reflecting_boundary(m)) for m in mid], bool)
        for m, out_p, out_i in ((~mask, absorbing_positions, absorbing_indices), (mask, reflecting_positions, reflecting_indices)):
            t = idx[m]
            used, first = np.unique(t.reshape(-1), return_index=True)
            order = used[np.argsort(first)]                      # vertices in order of first use, as the reference's maps fill
            remap = np.full(len(p), -1, np.int64)
            remap[order] = np.arange(len(order))
            _refill(out_p, list(p[order]))
            _refill(out_i, list(remap[t].astype(np.int32)))

    @staticmethod
    def compute_signed_volume(positions, indices, *, dim):
        p, idx = _points(positions, dim).astype(np.float64), _points(indices, dim, np.int32)
        if dim == 2:
            a, b = p[idx[:, 0]], p[idx[:, 1]]
            return float(0.5*np.sum(a[:, 0]*b[:, 1] - a[:, 1]*b[:, 0]))
        a, b, c = p[idx[:, 0]], p[idx[:, 1]], p[idx[:, 2]]
        return float(np.sum(np.einsum("ij,ij->i", a, np.cross(b, c)))/6.0)

    @staticmethod
    def FcpwDirichletBoundaryHandler(*, dim):
        _check_dim(dim)
        return _DirichletHandler(dim)

    @staticmethod
    def FcpwNeumannBoundaryHandler(*, dim):
        _check_dim(dim)
        return _NeumannHandler(dim)

    @staticmethod
    def FcpwRobinBoundaryHandler(*, dim):
        _check_dim(dim)
        return _RobinHandler(dim)

    @staticmethod
    def get_ignore_candidate_silhouette_callback(solve_double_sided=False, silhouette_precision=1e-3):
        return IgnoreCandidateSilhouette(solve_double_sided, silhouette_precision)

    @staticmethod
    def get_branch_traversal_weight_callback(min_radial_dist=1e-2):
        return BranchTraversalWeight(min_radial_dist)

    @staticmethod
    def populate_geometric_queries_for_dirichlet_boundary(handler, geometric_queries, *, dim):
        if not isinstance(handler, _DirichletHandler):
            raise TypeError("SDF grids are out of scope: pass a FcpwDirichletBoundaryHandler")
        geometric_queries._absorbing = handler
        geometric_queries.has_non_empty_absorbing_boundary = len(handler.indices) > 0
        geometric_queries._invalidate()

    @staticmethod
    def populate_geometric_queries_for_neumann_boundary(handler, branch_traversal_weight, geometric_queries, *, dim):
        geometric_queries._reflecting = handler
        geometric_queries.has_non_empty_reflecting_boundary = len(handler.indices) > 0
        geometric_queries._invalidate()

    @staticmethod
    def populate_geometric_queries_for_robin_boundary(handler, branch_traversal_weight, geometric_queries, *, dim):
        geometric_queries._reflecting = handler
        geometric_queries.has_non_empty_reflecting_boundary = len(handler.indices) > 0
        handler._queries.append(geometric_queries)
        geometric_queries._invalidate()

    @staticmethod
    def compute_dist_to_boundary(geometric_queries, solve_locations, dist_to_absorbing_boundary, dist_to_reflecting_boundary, *, dim):
        """the batched distances of the demo (zombie_bindings.cpp: computeDistToBoundary): one zb_closest_point per boundary"""
        q = geometric_queries
        x = _points(solve_locations, dim)
        big = np.full(len(x), abi.FLT_MAX, np.float32)
        da = q.scene().closest_point(abi.ABSORBING, x, want_normal=False)[1] if q.has_non_empty_absorbing_boundary and len(x) else big
        dr = q.scene().closest_point(abi.REFLECTING, x, want_normal=False)[1] if q.has_non_empty_reflecting_boundary and len(x) else big
        _refill(dist_to_absorbing_boundary, [float(v) for v in da])
        _refill(dist_to_reflecting_boundary, [float(v) for v in dr])

    @staticmethod
    def _dense(data, args, dim, with_aligned, channels=1):
        # (data, shape, min, max[, interp]) or (data, aligned data, shape, min, max[, interp])
        if with_aligned and len(args) >= 4 and np.asarray(args[1]).size == dim and np.asarray(args[0]).size != dim:
            aligned, shape, lo, hi = args[0], args[1], args[2], args[3]
            interp = bool(args[4]) if len(args) > 4 else False
            return DenseGridCallback(data, shape, lo, hi, dim, interp, aligned, channels)
        shape, lo, hi = args[0], args[1], args[2]
        interp = bool(args[3]) if len(args) > 3 else False
        return DenseGridCallback(data, shape, lo, hi, dim, interp, None, channels)

    @staticmethod
    def get_dense_grid_source_callback(data, *args, dim, channels=1):
        _check_dim(dim); _check_channels(channels)
        return Utils._dense(data, args, dim, False, channels)

    @staticmethod
    def get_dense_grid_dirichlet_callback(data, *args, dim, channels=1):
        _check_dim(dim); _check_channels(channels)
        return Utils._dense(data, args, dim, True, channels)

    @staticmethod
    def get_dense_grid_robin_callback(data, *args, dim, channels=1):
        _check_dim(dim); _check_channels(channels)
        return Utils._dense(data, args, dim, True, channels)

    @staticmethod
    def get_dense_grid_robin_coefficient_callback(data, *args, dim):
        _check_dim(dim)
        return Utils._dense(data, args, dim, False)

    @staticmethod
    def get_dense_grid_indicator_callback(data, shape, grid_min, grid_max, *, dim):
        _check_dim(dim)
        g = DenseGridCallback(np.asarray(data, np.float32), shape, grid_min, grid_max, dim, False)
        return lambda x: g(x) > 0.0

    class ProgressBar:
        def __init__(self, total_work=0, display_width=80):
            self.total_work = total_work

        def report(self, new_work_completed, thread_id=0):
            pass

        def finish(self):
            pass

    @staticmethod
    def get_report_progress_callback(progress_bar):
        return progress_bar.report


class Samplers:
    """The samplers of the reference (variance_reduction/{boundary,domain}_samplers.h) live in zombie_b200.bvc; these factories
    keep the reference's names."""

    @staticmethod
    def create_uniform_domain_sampler(geometric_queries, inside_solve_region, solve_region_min, solve_region_max, solve_region_volume, *,
                                      dim, channels=1):
        from . import bvc
        _check_dim(dim); _check_channels(channels)
        inside = inside_solve_region or geometric_queries.inside_domain_batch
        return bvc.UniformDomainSampler(inside, solve_region_min, solve_region_max, solve_region_volume)

    @staticmethod
    def create_uniform_line_segment_boundary_sampler(positions, indices, geometric_queries=None, inside_solve_region=None, *, dim, channels=1):
        from . import bvc
        if dim != 2:
            raise ValueError("create_uniform_line_segment_boundary_sampler is only supported for dim=2, got %r" % (dim,))
        return bvc.UniformBoundarySampler(_points(positions, 2), _points(indices, 2, np.int32))

    @staticmethod
    def create_uniform_triangle_boundary_sampler(positions, indices, geometric_queries=None, inside_solve_region=None, *, dim, channels=1):
        from . import bvc
        if dim != 3:
            raise ValueError("create_uniform_triangle_boundary_sampler is only supported for dim=3, got %r" % (dim,))
        return bvc.UniformBoundarySampler(_points(positions, 3), _points(indices, 3, np.int32))


__all__ = ["BoolList", "IntList", "UintList", "FloatList", "Float2List", "Float3List", "Int2List", "Int3List", "FloatNList", "IntNList",
           "convert_list_to_numpy_array", "Core", "Solvers", "Samplers", "Utils"]
