"""TEST INFRASTRUCTURE: ctypes binding of oracle/_ref/libzombie_ref.so — the unmodified reference
(zombie + FCPW + pcg32 + oneTBB) compiled by oracle/Makefile. Imported only by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs; never by zombie_b200/.
"""
import ctypes as C
import os

import numpy as np

from zombie_b200 import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_ref", "libzombie_ref.so")
_lib = None


def available():
    return os.path.exists(_LIB_PATH)


def cpu_supports_v3():
    """AVX2 + FMA + BMI2 on this host (what oracle/_ref/libzombie_ref_v3.so was compiled for)"""
    try:
        flags = open("/proc/cpuinfo").read()
    except OSError:
        return False
    return all((" " + f + " ") in flags.replace("\n", " ") for f in ("avx2", "fma", "bmi2"))


def select_build(name):
    """Timed CPU arm only: "" = the parity build (-O3 -DNDEBUG), "v3" = the same sources with -march=x86-64-v3. Scenes made
    before the switch keep their library."""
    global _lib, _LIB_PATH
    path = os.path.join(_HERE, "_ref", "libzombie_ref%s.so" % ("_" + name if name else ""))
    if not os.path.exists(path):
        return False
    if path != _LIB_PATH:
        _LIB_PATH, _lib = path, None
    return True


def lib():
    global _lib
    if _lib is None:
        C.CDLL(os.path.join(_HERE, "_ref", "libtbb.so.12"), mode=C.RTLD_GLOBAL)
        _lib = C.CDLL(_LIB_PATH)
        _lib.zref_scene_create.restype = C.c_void_p
        _lib.zref_scene_create.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                           C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                           C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_float,
                                           C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        _lib.zref_scene_destroy.argtypes = [C.c_void_p]
        _lib.zref_closest_point.argtypes = [C.c_void_p, C.c_int, C.c_int] + [C.c_void_p]*6 + [C.c_int]
        _lib.zref_intersect_ray.argtypes = [C.c_void_p, C.c_int, C.c_int] + [C.c_void_p]*7 + [C.c_int]
        _lib.zref_occluded.argtypes = [C.c_void_p, C.c_int, C.c_int] + [C.c_void_p]*3 + [C.c_int]
        _lib.zref_count_ray_hits.argtypes = [C.c_void_p, C.c_int, C.c_int] + [C.c_void_p]*4 + [C.c_int]
        _lib.zref_inside_domain.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        _lib.zref_sample_sphere.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p]*7 + [C.c_int]
        _lib.zref_star_radius.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float,
                                          C.c_void_p, C.c_void_p, C.c_int]
        _lib.zref_solve.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_int, C.c_void_p]
        _lib.zref_bvc_splat.argtypes = [C.c_int, C.c_float, C.c_void_p] + [C.c_float]*5 + [C.c_int] + [C.c_void_p]*9
        _lib.zref_kelvin_transform_points.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        _lib.zref_kelvin_robin_coefficients.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int] + [C.c_void_p]*5
        _lib.zref_kelvin_back_transform.argtypes = [C.c_int, C.c_void_p, C.c_int] + [C.c_void_p]*5
        _lib.zref_rws_solve.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_void_p,
                                        C.c_int, C.c_void_p, C.c_void_p]
        _lib.zref_pcg32.argtypes = [C.c_uint64, C.c_uint64, C.c_int, C.c_void_p]
        _lib.zref_pcg32_demo.argtypes = [C.c_uint64, C.c_uint64, C.c_int] + [C.c_void_p]*4
        _lib.zref_pcg32_floats.argtypes = [C.c_uint64, C.c_uint64, C.c_int, C.c_void_p]
    return _lib


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def hardware_threads():
    return lib().zref_hardware_threads()


def pcg32(initstate, initseq, n):
    out = np.zeros(n, np.uint32)
    lib().zref_pcg32(initstate, initseq, n, abi.ptr(out))
    return out


def pcg32_demo(initstate, initseq, rounds):
    """the sequence deps/pcg32/pcg32-demo.cpp prints: (u32 [r,6], coins [r,65], rolls [r,33], cards [r,52])"""
    u = np.zeros((rounds, 6), np.uint32)
    coins = np.zeros((rounds, 65), np.uint8)
    rolls = np.zeros((rounds, 33), np.uint8)
    cards = np.zeros((rounds, 52), np.uint8)
    lib().zref_pcg32_demo(initstate, initseq, rounds, abi.ptr(u), abi.ptr(coins), abi.ptr(rolls), abi.ptr(cards))
    return u, coins, rolls, cards


def pcg32_floats(initstate, initseq, n):
    out = np.zeros(n, np.float32)
    lib().zref_pcg32_floats(initstate, initseq, n, abi.ptr(out))
    return out


def bvc_splat(dim, absorption_coeff, samples, eval_pts, radius_clamp=1e-3, kernel_regularization=0.0,
              robin_cutoff=abi.FLT_MAX, cutoff_a=0.0, cutoff_r=0.0):
    """bvc::BoundaryValueCaching::splat of the reference on abi.HostBvcSamples / abi.HostBvcEvalPoints.
    Returns (solution [n], gradient [n, dim], bucketCount [n, 5], bucketMean [n, 5, 1 + dim], seconds)."""
    n = eval_pts.n
    sol = np.zeros(n, np.float32)
    grad = np.zeros((n, dim), np.float32)
    cnt = np.zeros((n, 5), np.int32)
    mean = np.zeros((n, 5, 1 + dim), np.float32)
    secs = C.c_double(0.0)
    ss = samples.as_struct()
    rc = lib().zref_bvc_splat(dim, absorption_coeff, C.byref(ss), radius_clamp, kernel_regularization, robin_cutoff,
                              cutoff_a, cutoff_r, n, abi.ptr(eval_pts.pt), abi.ptr(eval_pts.type),
                              abi.ptr(eval_pts.distToAbsorbingBoundary), abi.ptr(eval_pts.distToReflectingBoundary),
                              abi.ptr(sol), abi.ptr(grad), abi.ptr(cnt), abi.ptr(mean), C.byref(secs))
    assert rc == 0
    return sol, grad, cnt, mean, secs.value


def kelvin_transform_points(x, origin):
    x, origin = _f32(x), _f32(origin)
    y = np.zeros_like(x)
    lib().zref_kelvin_transform_points(x.shape[1], abi.ptr(origin), len(x), abi.ptr(x), abi.ptr(y))
    return y


def kelvin_robin_coefficients(transformed_points, indices, min_coeff, max_coeff, origin):
    p, mn, mx, origin = _f32(transformed_points), _f32(min_coeff), _f32(max_coeff), _f32(origin)
    idx = np.ascontiguousarray(indices, np.int32)
    omn, omx = np.zeros(len(idx), np.float32), np.zeros(len(idx), np.float32)
    lib().zref_kelvin_robin_coefficients(p.shape[1], abi.ptr(origin), len(p), abi.ptr(p), len(idx), abi.ptr(idx),
                                         abi.ptr(mn), abi.ptr(mx), abi.ptr(omn), abi.ptr(omx))
    return omn, omx


def kelvin_back_transform(V, dV, y, origin):
    y, V, dV, origin = _f32(y), _f32(V), _f32(dV), _f32(origin)
    u, du = np.zeros(len(y), np.float32), np.zeros_like(y)
    lib().zref_kelvin_back_transform(y.shape[1], abi.ptr(origin), len(y), abi.ptr(y), abi.ptr(V), abi.ptr(dV), abi.ptr(u), abi.ptr(du))
    return u, du


class RefScene:
    """Fcpw{Dirichlet,Neumann,Robin}BoundaryHandler + populated GeometricQueries of the reference."""

    def __init__(self, dim, absorbing=None, reflecting=None, reflecting_kind=abi.REFLECTING_NONE,
                 min_robin=None, max_robin=None, solve_double_sided=False, silhouette_precision=1e-3,
                 watertight=True, domain_min=None, domain_max=None, use_baseline=False):
        self.dim = dim
        ap, ai = absorbing if absorbing is not None else (np.zeros((0, dim), np.float32), np.zeros((0, dim), np.int32))
        rp, ri = reflecting if reflecting is not None else (np.zeros((0, dim), np.float32), np.zeros((0, dim), np.int32))
        ap, rp = _f32(ap), _f32(rp)
        ai = np.ascontiguousarray(ai, np.int32)
        ri = np.ascontiguousarray(ri, np.int32)
        allp = np.concatenate([ap, rp], axis=0)
        if domain_min is None:
            domain_min = allp.min(axis=0) - 1e-3
        if domain_max is None:
            domain_max = allp.max(axis=0) + 1e-3
        self.domain_min, self.domain_max = _f32(domain_min), _f32(domain_max)
        mn = _f32(min_robin) if min_robin is not None else np.zeros(len(ri), np.float32)
        mx = _f32(max_robin) if max_robin is not None else np.zeros(len(ri), np.float32)
        self._lib = lib()
        self._h = self._lib.zref_scene_create(dim, abi.ptr(ap), len(ap), abi.ptr(ai), len(ai),
                                          abi.ptr(rp), len(rp), abi.ptr(ri), len(ri),
                                          reflecting_kind, abi.ptr(mn), abi.ptr(mx),
                                          int(solve_double_sided), silhouette_precision,
                                          int(watertight), abi.ptr(self.domain_min), abi.ptr(self.domain_max),
                                          int(use_baseline))

    def close(self):
        if self._h:
            self._lib.zref_scene_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def closest_point(self, which, x, max_radius=None, want_normal=True, threads=0):
        x = _f32(x)
        n = len(x)
        prim = np.zeros(n, np.int32)
        dist = np.zeros(n, np.float32)
        pt = np.zeros((n, self.dim), np.float32)
        nrm = np.zeros((n, self.dim), np.float32) if want_normal else None
        rc = self._lib.zref_closest_point(self._h, which, n, abi.ptr(x), abi.ptr(_f32(max_radius)),
                                      abi.ptr(prim), abi.ptr(dist), abi.ptr(pt), abi.ptr(nrm), threads)
        assert rc == 0
        return prim, dist, pt, nrm

    def intersect_ray(self, which, origin, direction, tmax=None, threads=0):
        o, d = _f32(origin), _f32(direction)
        n = len(o)
        prim = np.zeros(n, np.int32)
        dist = np.zeros(n, np.float32)
        pt = np.zeros((n, self.dim), np.float32)
        nrm = np.zeros((n, self.dim), np.float32)
        rc = self._lib.zref_intersect_ray(self._h, which, n, abi.ptr(o), abi.ptr(d), abi.ptr(_f32(tmax)),
                                      abi.ptr(prim), abi.ptr(dist), abi.ptr(pt), abi.ptr(nrm), threads)
        assert rc == 0
        return prim, dist, pt, nrm

    def occluded(self, which, a, b, threads=0):
        a, b = _f32(a), _f32(b)
        out = np.zeros(len(a), np.uint8)
        rc = self._lib.zref_occluded(self._h, which, len(a), abi.ptr(a), abi.ptr(b), abi.ptr(out), threads)
        assert rc == 0
        return out

    def count_ray_hits(self, which, origin, direction, tmax=None, threads=0):
        o, d = _f32(origin), _f32(direction)
        t = None if tmax is None else _f32(tmax)
        out = np.zeros(len(o), np.int32)
        rc = self._lib.zref_count_ray_hits(self._h, which, len(o), abi.ptr(o), abi.ptr(d), abi.ptr(t) if t is not None else None,
                                           abi.ptr(out), threads)
        assert rc == 0
        return out

    def inside_domain(self, x, threads=0):
        x = _f32(x)
        out = np.zeros(len(x), np.uint8)
        rc = self._lib.zref_inside_domain(self._h, len(x), abi.ptr(x), abi.ptr(out), threads)
        assert rc == 0
        return out.astype(bool)

    def sample_sphere(self, x, radius, rand_nums, threads=0):
        x, radius, u = _f32(x), _f32(radius), _f32(rand_nums)
        n = len(x)
        prim = np.zeros(n, np.int32)
        pt = np.zeros((n, self.dim), np.float32)
        nrm = np.zeros((n, self.dim), np.float32)
        pdf = np.zeros(n, np.float32)
        rc = self._lib.zref_sample_sphere(self._h, n, abi.ptr(x), abi.ptr(radius), abi.ptr(u),
                                      abi.ptr(prim), abi.ptr(pt), abi.ptr(nrm), abi.ptr(pdf), threads)
        assert rc == 0
        return prim, pt, nrm, pdf

    def star_radius(self, x, min_radius, max_radius, precision=1e-3, flip=None, threads=0):
        x, mn, mx = _f32(x), _f32(min_radius), _f32(max_radius)
        out = np.zeros(len(x), np.float32)
        fl = None if flip is None else np.ascontiguousarray(flip, np.uint8)
        rc = self._lib.zref_star_radius(self._h, len(x), abi.ptr(x), abi.ptr(mn), abi.ptr(mx), precision,
                                    abi.ptr(fl), abi.ptr(out), threads)
        assert rc == 0
        return out

    def rws_solve(self, pde, settings, samples, eval_pts, normal_offset, radius_clamp=0.0, kernel_regularization=0.0, threads=0):
        """rws::ReverseWalkOnStarsSolver::solve on abi.HostRwsSamples / abi.HostRwsEvalPoints (totals accumulate).
        Returns (group counts [5], seconds)."""
        counts = np.zeros(5, np.int32)
        secs = C.c_double(0.0)
        ss, es = samples.as_struct(), eval_pts.as_struct()
        rc = self._lib.zref_rws_solve(self._h, C.byref(pde), C.byref(settings), C.byref(ss), normal_offset, radius_clamp,
                                  kernel_regularization, C.byref(es), threads, abi.ptr(counts), C.byref(secs))
        assert rc == 0
        return counts, secs.value

    def solve(self, estimator, pde, settings, pts, stats, threads=0):
        """WalkOnSpheres/WalkOnStars::solve; threads: 0 = TBB default arena, 1 = runSingleThreaded.
        Returns (walk steps taken, seconds inside solve())."""
        steps = C.c_int64(0)
        secs = C.c_double(0.0)
        ps, ss = pts.as_struct(), stats.as_struct()
        rc = self._lib.zref_solve(self._h, estimator, C.byref(pde), C.byref(settings), C.byref(ps), C.byref(ss),
                              C.byref(steps), threads, C.byref(secs))
        assert rc == 0
        return steps.value, secs.value
