"""ctypes mirrors of the C ABI structs declared in include/zombie_b200.h."""
import ctypes as C

import numpy as np

ZB_OK = 0

# enum zb_sample_type / zb_estimation_quantity / zb_reflecting_kind / zb_estimator / zb_boundary_id
IN_DOMAIN, ON_ABSORBING_BOUNDARY, ON_REFLECTING_BOUNDARY = 0, 1, 2
SOLUTION, SOLUTION_AND_GRADIENT, SKIP = 0, 1, 2
REFLECTING_NONE, REFLECTING_NEUMANN, REFLECTING_ROBIN = 0, 1, 2
BVH_BUILDER_HOST_SAH, BVH_BUILDER_DEVICE = 0, 1   # zb_bvh_builder
WALK_ON_SPHERES, WALK_ON_STARS = 0, 1
ABSORBING, REFLECTING = 0, 1
PDE_MANUFACTURED, PDE_CONSTANT = 0, 1
(SOL_CONSTANT, SOL_EXP_DIR, SOL_QUADRATIC, SOL_INV_DIST, SOL_SIN_EXP_2D, SOL_CUBIC_2D,
 SOL_LINEAR, SOL_TRIG_3D) = range(8)

FLT_MAX = float(np.finfo(np.float32).max)


class WalkSettings(C.Structure):
    """zb_walk_settings == zombie::WalkSettings (point_estimation/common.h:19-91)."""
    _fields_ = [("epsilonShellForAbsorbingBoundary", C.c_float),
                ("epsilonShellForReflectingBoundary", C.c_float),
                ("silhouettePrecision", C.c_float),
                ("russianRouletteThreshold", C.c_float),
                ("splittingThreshold", C.c_float),
                ("maxWalkLength", C.c_int32),
                ("stepsBeforeApplyingTikhonov", C.c_int32),
                ("stepsBeforeUsingMaximalSpheres", C.c_int32),
                ("solveDoubleSided", C.c_int32),
                ("useGradientControlVariates", C.c_int32),
                ("useGradientAntitheticVariates", C.c_int32),
                ("useCosineSamplingForDerivatives", C.c_int32),
                ("ignoreAbsorbingBoundaryContribution", C.c_int32),
                ("ignoreReflectingBoundaryContribution", C.c_int32),
                ("ignoreSourceContribution", C.c_int32),
                ("printLogs", C.c_int32)]

    @classmethod
    def default(cls, epsilonShellForAbsorbingBoundary=1e-3, epsilonShellForReflectingBoundary=1e-3,
                maxWalkLength=1024, solveDoubleSided=False, **kw):
        """The short WalkSettings constructor (common.h:21-40)."""
        s = cls(epsilonShellForAbsorbingBoundary, epsilonShellForReflectingBoundary, 1e-3, 0.0, FLT_MAX,
                maxWalkLength, 0, maxWalkLength, int(solveDoubleSided), 1, 1, 0, 0, 0, 0, 0)
        for k, v in kw.items():
            if not hasattr(s, k):
                raise AttributeError(k)
            setattr(s, k, v)
        return s


GRID_NEAREST, GRID_LINEAR = 0, 1


class GridView(C.Structure):
    """zb_grid_view: dense regular samples of one PDE callback; data is a device pointer for the product (native.Grid),
    a host pointer for the test oracle."""
    _fields_ = [("data", C.c_void_p), ("shape", C.c_int32 * 3), ("interp", C.c_int32),
                ("origin", C.c_float * 3), ("invCell", C.c_float * 3)]


class HostGrid:
    """Samples of a callback on a regular lattice over [lo, hi] (both ends included): values[k, j, i] at
    lo + (i, j, k)*(hi - lo)/(shape - 1); a 2D grid has shape (1, ny, nx)."""

    def __init__(self, values, lo, hi, interp=GRID_LINEAR):
        v = np.ascontiguousarray(values, np.float32)
        if v.ndim == 2:
            v = v[None]
        self.values = v
        self.interp = interp
        self.lo = [float(t) for t in lo] + [0.0]*(3 - len(lo))
        self.hi = [float(t) for t in hi] + [0.0]*(3 - len(hi))

    @classmethod
    def sample(cls, fn, lo, hi, shape, interp=GRID_LINEAR):
        """tabulates fn(points [n, dim]) -> [n] on the lattice; shape = (nx, ny[, nz])"""
        axes = [np.linspace(lo[k], hi[k], shape[k], dtype=np.float64) for k in range(len(shape))]
        mesh = np.meshgrid(*axes, indexing="ij")                   # [nx, ny(, nz)]
        pts = np.stack([m.ravel() for m in mesh], axis=1).astype(np.float32)
        vals = np.asarray(fn(pts), np.float32).reshape([len(a) for a in axes])
        vals = vals.transpose(*reversed(range(vals.ndim)))          # -> [(nz,) ny, nx]: x fastest
        return cls(vals, lo, hi, interp)

    def view(self, data_ptr=None):
        g = GridView()
        g.data = self.values.ctypes.data if data_ptr is None else data_ptr
        nz, ny, nx = self.values.shape
        for k, n in enumerate((nx, ny, nz)):
            g.shape[k] = n
            g.origin[k] = self.lo[k]
            g.invCell[k] = (n - 1)/(self.hi[k] - self.lo[k]) if n > 1 else 0.0
        g.interp = self.interp
        return g


class PdeDesc(C.Structure):
    """zb_pde_desc: data-described zombie::PDE (core/pde.h:20-45)."""
    _fields_ = [("absorptionCoeff", C.c_float),
                ("isSourceConstant", C.c_int32),
                ("areRobinConditionsPureNeumann", C.c_int32),
                ("areRobinCoeffsNonnegative", C.c_int32),
                ("mode", C.c_int32),
                ("family", C.c_int32),
                ("params", C.c_float * 16),
                ("robinCoeff", C.c_float),
                ("sourceValue", C.c_float),
                ("dirichletValue", C.c_float),
                ("robinValue", C.c_float),
                ("kelvin", C.c_int32),
                ("kelvinOrigin", C.c_float * 3),
                ("sourceGrid", GridView), ("dirichletGrid", GridView), ("robinGrid", GridView),
                ("robinAlignedGrid", GridView), ("robinCoeffGrid", GridView),
                ("dirichletAlignedGrid", GridView), ("dirichletAlignedValue", C.c_float), ("hasDirichletAlignedValue", C.c_int32)]

    GRID_FIELDS = ("sourceGrid", "dirichletGrid", "robinGrid", "robinAlignedGrid", "robinCoeffGrid", "dirichletAlignedGrid")

    def set_grids(self, **grids):
        """attaches HostGrid objects (source=, dirichlet=, dirichletAligned=, robin=, robinAligned=, robinCoeff=). The structure then carries
        HOST pointers (what the test oracle reads); native.Scene.solve uploads the grids and passes device pointers."""
        self.__dict__.setdefault("host_grids", {})
        for k, g in grids.items():
            name = k + "Grid"
            assert name in self.GRID_FIELDS
            self.host_grids[name] = g
            setattr(self, name, g.view())
        return self

    @classmethod
    def manufactured(cls, family, params, absorptionCoeff=0.0, robinCoeff=0.0, isSourceConstant=False):
        d = cls()
        d.absorptionCoeff = absorptionCoeff
        d.isSourceConstant = int(isSourceConstant)
        d.areRobinConditionsPureNeumann = int(robinCoeff == 0.0)
        d.areRobinCoeffsNonnegative = int(robinCoeff >= 0.0)
        d.mode = PDE_MANUFACTURED
        d.family = family
        for i, p in enumerate(params):
            d.params[i] = p
        d.robinCoeff = robinCoeff
        return d


class SamplePoints(C.Structure):
    _fields_ = [("n", C.c_int32),
                ("pt", C.c_void_p), ("normal", C.c_void_p), ("directionForDerivative", C.c_void_p),
                ("type", C.c_void_p), ("estimationQuantity", C.c_void_p),
                ("distToAbsorbingBoundary", C.c_void_p), ("distToReflectingBoundary", C.c_void_p),
                ("estimateBoundaryNormalAligned", C.c_void_p), ("nWalks", C.c_void_p),
                ("rngState", C.c_void_p), ("rngInc", C.c_void_p), ("firstSphereRadius", C.c_void_p)]


class Statistics(C.Structure):
    _fields_ = [("solutionMean", C.c_void_p), ("solutionM2", C.c_void_p),
                ("gradientMean", C.c_void_p), ("gradientM2", C.c_void_p),
                ("totalFirstSourceContribution", C.c_void_p), ("totalDerivativeContribution", C.c_void_p),
                ("nSolutionEstimates", C.c_void_p), ("nGradientEstimates", C.c_void_p),
                ("totalWalkLength", C.c_void_p), ("totalSplits", C.c_void_p)]


def ptr(a):
    """void* of a C-contiguous numpy array (None -> NULL)."""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


class HostSamplePoints:
    """Owns the numpy arrays behind a zb_sample_points (SoA of zombie::SamplePoint)."""

    def __init__(self, pt, distToAbsorbingBoundary, distToReflectingBoundary=None, nWalks=1,
                 estimationQuantity=SOLUTION, type=IN_DOMAIN, normal=None, seed=0xC0FFEE, streams=None):
        from .procedural import pcg32_seed
        pt = np.ascontiguousarray(pt, dtype=np.float32)
        n, dim = pt.shape
        self.n, self.dim = n, dim
        self.pt = pt
        self.normal = np.zeros((n, dim), np.float32) if normal is None else np.ascontiguousarray(normal, np.float32)
        self.directionForDerivative = np.zeros((n, dim), np.float32)
        self.directionForDerivative[:, 0] = 1.0
        self.type = np.full(n, type, np.int32) if np.isscalar(type) else np.ascontiguousarray(type, np.int32)
        self.estimationQuantity = (np.full(n, estimationQuantity, np.int32) if np.isscalar(estimationQuantity)
                                   else np.ascontiguousarray(estimationQuantity, np.int32))
        self.distToAbsorbingBoundary = np.ascontiguousarray(distToAbsorbingBoundary, np.float32)
        self.distToReflectingBoundary = (np.full(n, FLT_MAX, np.float32) if distToReflectingBoundary is None
                                         else np.ascontiguousarray(distToReflectingBoundary, np.float32))
        self.estimateBoundaryNormalAligned = np.zeros(n, np.uint8)
        self.nWalks = np.full(n, nWalks, np.int32) if np.isscalar(nWalks) else np.ascontiguousarray(nWalks, np.int32)
        if streams is None:
            streams = np.arange(n, dtype=np.uint64)
        self.rngState, self.rngInc = pcg32_seed(np.full(n, seed, np.uint64), np.asarray(streams, np.uint64))
        self.rngState = np.ascontiguousarray(self.rngState)
        self.rngInc = np.ascontiguousarray(self.rngInc)
        self.firstSphereRadius = np.zeros(n, np.float32)

    def as_struct(self):
        return SamplePoints(self.n, ptr(self.pt), ptr(self.normal), ptr(self.directionForDerivative),
                            ptr(self.type), ptr(self.estimationQuantity), ptr(self.distToAbsorbingBoundary),
                            ptr(self.distToReflectingBoundary), ptr(self.estimateBoundaryNormalAligned),
                            ptr(self.nWalks), ptr(self.rngState), ptr(self.rngInc), ptr(self.firstSphereRadius))

    def subset(self, idx):
        o = object.__new__(HostSamplePoints)
        o.n, o.dim = len(idx), self.dim
        for k in ("pt", "normal", "directionForDerivative", "type", "estimationQuantity",
                  "distToAbsorbingBoundary", "distToReflectingBoundary", "estimateBoundaryNormalAligned",
                  "nWalks", "rngState", "rngInc", "firstSphereRadius"):
            setattr(o, k, np.ascontiguousarray(getattr(self, k)[idx]))
        return o


class HostStatistics:
    """Owns the numpy arrays behind a zb_statistics (SoA of zombie::SampleStatistics)."""

    def __init__(self, n, dim):
        self.n, self.dim = n, dim
        self.solutionMean = np.zeros(n, np.float32)
        self.solutionM2 = np.zeros(n, np.float32)
        self.gradientMean = np.zeros((n, dim), np.float32)
        self.gradientM2 = np.zeros((n, dim), np.float32)
        self.totalFirstSourceContribution = np.zeros(n, np.float32)
        self.totalDerivativeContribution = np.zeros(n, np.float32)
        self.nSolutionEstimates = np.zeros(n, np.int32)
        self.nGradientEstimates = np.zeros(n, np.int32)
        self.totalWalkLength = np.zeros(n, np.int32)
        self.totalSplits = np.zeros(n, np.int32)

    def as_struct(self):
        return Statistics(ptr(self.solutionMean), ptr(self.solutionM2), ptr(self.gradientMean), ptr(self.gradientM2),
                          ptr(self.totalFirstSourceContribution), ptr(self.totalDerivativeContribution),
                          ptr(self.nSolutionEstimates), ptr(self.nGradientEstimates),
                          ptr(self.totalWalkLength), ptr(self.totalSplits))

    # SampleStatistics getters (common.h:313-377)
    def solution_variance(self):
        return self.solutionM2 / np.maximum(1, self.nSolutionEstimates - 1)

    def gradient_variance(self):
        return self.gradientM2 / np.maximum(1, self.nGradientEstimates - 1)[:, None]

    def mean_walk_length(self):
        return self.totalWalkLength / np.maximum(1, self.nSolutionEstimates)


BVC_BUCKETS = 5  # absorbing, absorbing normal-aligned, reflecting, reflecting normal-aligned, source


class BvcSamples(C.Structure):
    """zb_bvc_samples: SoA of the zombie::SamplePoint fields BoundaryValueCaching::splat reads."""
    _fields_ = [("n", C.c_int32), ("pt", C.c_void_p), ("normal", C.c_void_p), ("type", C.c_void_p),
                ("estimateBoundaryNormalAligned", C.c_void_p), ("pdf", C.c_void_p), ("solution", C.c_void_p),
                ("normalDerivative", C.c_void_p), ("contribution", C.c_void_p), ("robinCoeff", C.c_void_p)]


class BvcEvalPoints(C.Structure):
    """zb_bvc_eval_points: SoA of bvc::EvaluationPoint with (count, sums) per statistics bucket."""
    _fields_ = [("n", C.c_int32), ("pt", C.c_void_p), ("type", C.c_void_p), ("distToAbsorbingBoundary", C.c_void_p),
                ("distToReflectingBoundary", C.c_void_p), ("bucketCount", C.c_void_p), ("bucketSum", C.c_void_p),
                ("solution", C.c_void_p), ("gradient", C.c_void_p)]


class HostBvcSamples:
    def __init__(self, pt, normal, type, pdf, solution, normalDerivative, contribution, robinCoeff, aligned=None):
        self.pt = np.ascontiguousarray(pt, np.float32)
        self.n, self.dim = self.pt.shape
        self.normal = np.ascontiguousarray(normal, np.float32)
        self.type = np.ascontiguousarray(type, np.int32)
        self.estimateBoundaryNormalAligned = (np.zeros(self.n, np.uint8) if aligned is None
                                              else np.ascontiguousarray(aligned, np.uint8))
        self.pdf = np.ascontiguousarray(pdf, np.float32)
        self.solution = np.ascontiguousarray(solution, np.float32)
        self.normalDerivative = np.ascontiguousarray(normalDerivative, np.float32)
        self.contribution = np.ascontiguousarray(contribution, np.float32)
        self.robinCoeff = np.ascontiguousarray(robinCoeff, np.float32)

    def as_struct(self):
        return BvcSamples(self.n, ptr(self.pt), ptr(self.normal), ptr(self.type), ptr(self.estimateBoundaryNormalAligned),
                          ptr(self.pdf), ptr(self.solution), ptr(self.normalDerivative), ptr(self.contribution),
                          ptr(self.robinCoeff))

    def subset(self, idx):
        return HostBvcSamples(self.pt[idx], self.normal[idx], self.type[idx], self.pdf[idx], self.solution[idx],
                              self.normalDerivative[idx], self.contribution[idx], self.robinCoeff[idx],
                              self.estimateBoundaryNormalAligned[idx])


class HostBvcEvalPoints:
    def __init__(self, pt, distToAbsorbingBoundary, distToReflectingBoundary, type=IN_DOMAIN):
        self.pt = np.ascontiguousarray(pt, np.float32)
        self.n, self.dim = self.pt.shape
        self.type = np.full(self.n, type, np.int32) if np.isscalar(type) else np.ascontiguousarray(type, np.int32)
        self.distToAbsorbingBoundary = np.ascontiguousarray(distToAbsorbingBoundary, np.float32)
        self.distToReflectingBoundary = np.ascontiguousarray(distToReflectingBoundary, np.float32)
        self.bucketCount = np.zeros((self.n, BVC_BUCKETS), np.int32)
        self.bucketSum = np.zeros((self.n, BVC_BUCKETS, 1 + self.dim), np.float32)
        self.solution = np.zeros(self.n, np.float32)
        self.gradient = np.zeros((self.n, self.dim), np.float32)

    def as_struct(self):
        return BvcEvalPoints(self.n, ptr(self.pt), ptr(self.type), ptr(self.distToAbsorbingBoundary),
                             ptr(self.distToReflectingBoundary), ptr(self.bucketCount), ptr(self.bucketSum),
                             ptr(self.solution), ptr(self.gradient))

    def bucket_mean(self):
        """per-bucket Welford means [n, 5, 1 + dim] (0 where a bucket is empty)"""
        return self.bucketSum/np.maximum(1, self.bucketCount)[:, :, None]


class RwsSamples(C.Structure):
    """zb_rws_samples: SoA of the zombie::SamplePoint fields a reverse walk reads."""
    _fields_ = [("n", C.c_int32), ("pt", C.c_void_p), ("normal", C.c_void_p), ("type", C.c_void_p),
                ("estimateBoundaryNormalAligned", C.c_void_p), ("pdf", C.c_void_p), ("distToAbsorbingBoundary", C.c_void_p),
                ("distToReflectingBoundary", C.c_void_p), ("rngState", C.c_void_p), ("rngInc", C.c_void_p)]


class RwsEvalPoints(C.Structure):
    """zb_rws_eval_points: SoA of rws::EvaluationPoint; totals = 6 running sums per point."""
    _fields_ = [("n", C.c_int32), ("pt", C.c_void_p), ("normal", C.c_void_p), ("type", C.c_void_p), ("totals", C.c_void_p)]


class HostRwsSamples:
    def __init__(self, pt, normal, type, pdf, distToAbsorbingBoundary, distToReflectingBoundary, aligned=None, seed=0xC0FFEE,
                 streams=None):
        from .procedural import pcg32_seed
        self.pt = np.ascontiguousarray(pt, np.float32)
        self.n, self.dim = self.pt.shape
        self.normal = np.ascontiguousarray(normal, np.float32)
        self.type = np.ascontiguousarray(type, np.int32)
        self.estimateBoundaryNormalAligned = np.zeros(self.n, np.uint8) if aligned is None else np.ascontiguousarray(aligned, np.uint8)
        self.pdf = np.ascontiguousarray(pdf, np.float32)
        self.distToAbsorbingBoundary = np.ascontiguousarray(distToAbsorbingBoundary, np.float32)
        self.distToReflectingBoundary = np.ascontiguousarray(distToReflectingBoundary, np.float32)
        if streams is None:
            streams = np.arange(self.n, dtype=np.uint64)
        st, inc = pcg32_seed(np.full(self.n, seed, np.uint64), np.asarray(streams, np.uint64))
        self.rngState, self.rngInc = np.ascontiguousarray(st), np.ascontiguousarray(inc)

    def as_struct(self):
        return RwsSamples(self.n, ptr(self.pt), ptr(self.normal), ptr(self.type), ptr(self.estimateBoundaryNormalAligned),
                          ptr(self.pdf), ptr(self.distToAbsorbingBoundary), ptr(self.distToReflectingBoundary),
                          ptr(self.rngState), ptr(self.rngInc))

    def group_counts(self):
        """sample counts per statistics group: absorbing, absorbing aligned, reflecting, reflecting aligned, domain"""
        al = self.estimateBoundaryNormalAligned != 0
        a, r = self.type == ON_ABSORBING_BOUNDARY, self.type == ON_REFLECTING_BOUNDARY
        return np.array([np.sum(a & ~al), np.sum(a & al), np.sum(r & ~al), np.sum(r & al), np.sum(self.type == IN_DOMAIN)], np.int64)


class HostRwsEvalPoints:
    def __init__(self, pt, type=IN_DOMAIN, normal=None):
        self.pt = np.ascontiguousarray(pt, np.float32)
        self.n, self.dim = self.pt.shape
        self.normal = np.zeros_like(self.pt) if normal is None else np.ascontiguousarray(normal, np.float32)
        self.type = np.full(self.n, type, np.int32) if np.isscalar(type) else np.ascontiguousarray(type, np.int32)
        self.totals = np.zeros((self.n, 6), np.float32)

    def as_struct(self):
        return RwsEvalPoints(self.n, ptr(self.pt), ptr(self.normal), ptr(self.type), ptr(self.totals))

    def estimated_solution(self, group_counts, dirichlet=None):
        """rws::EvaluationPoint::getEstimatedSolution (reverse_walk_solver.h:129-166); points on the absorbing boundary
        take `dirichlet` (the solver assigns pde.dirichlet there, :324-329)"""
        t = self.totals.astype(np.float64)
        n = np.asarray(group_counts, np.float64)
        sol = np.zeros(self.n)
        if n[0] > 0:
            sol += np.where(t[:, 0] > 0.0, t[:, 1]/np.where(t[:, 0] > 0.0, t[:, 0], 1.0), t[:, 1]/n[0])
        for k, col in ((1, 2), (2, 3), (3, 4), (4, 5)):
            if n[k] > 0:
                sol += t[:, col]/n[k]
        on_abs = self.type == ON_ABSORBING_BOUNDARY
        if dirichlet is not None:
            sol = np.where(on_abs, dirichlet, sol)
        return sol.astype(np.float32)
