"""ctypes binding of zombie_b200/libzombie_b200.so (the C ABI declared in include/zombie_b200.h).

There is no CPU fallback: loading fails loudly when the CUDA library has not been built, and every
call raises ZombieError when the library reports an error (including "no CUDA device").
"""
import ctypes as C
import weakref
import os

import numpy as np

from . import abi
from .abi import FLT_MAX

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ZOMBIE_B200_LIB", os.path.join(_HERE, "libzombie_b200.so"))
_lib = None

# every symbol include/zombie_b200.h declares
SYMBOLS = ["zb_last_error", "zb_abi_version", "zb_device_count", "zb_context_create", "zb_context_destroy",
           "zb_context_set_stream", "zb_context_synchronize", "zb_context_device", "zb_context_sm_count",
           "zb_context_last_solve_profile",
           "zb_context_set_bvh_builder", "zb_scene_update_vertices", "zb_scene_update_vertices_dev", "zb_scene_build_profile",
           "zb_scene_validate", "zb_scene_download_records",
           "zb_host_bvh_check", "zb_scene_create", "zb_scene_destroy", "zb_scene_update_robin_coefficients", "zb_scene_update_robin_coefficients_dev", "zb_scene_stats",
           "zb_closest_point", "zb_intersect_ray", "zb_occluded", "zb_count_ray_hits", "zb_inside_domain", "zb_sample_sphere", "zb_star_radius",
           "zb_solve", "zb_solve_dev", "zb_solve_channels", "zb_solve_channels_dev", "zb_bvc_splat", "zb_bvc_splat_dev", "zb_context_last_splat_ms", "zb_rws_solve",
           "zb_grid_create", "zb_grid_data", "zb_grid_destroy",
           "zb_pcg32_kat", "zb_pcg32_demo", "zb_pcg32_floats"]


class ZombieError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ZombieError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(zombie_b200 has no CPU fallback)" % LIB_PATH)
        l = C.CDLL(LIB_PATH)
        l.zb_last_error.restype = C.c_char_p
        vp, i, f = C.c_void_p, C.c_int, C.c_float
        l.zb_context_create.argtypes = [i, C.POINTER(vp)]
        l.zb_context_destroy.argtypes = [vp]
        l.zb_context_destroy.restype = None
        l.zb_context_set_stream.argtypes = [vp, vp]
        l.zb_context_synchronize.argtypes = [vp]
        l.zb_context_device.argtypes = [vp]
        l.zb_context_sm_count.argtypes = [vp]
        l.zb_context_last_solve_profile.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float),
                                                    C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.POINTER(C.c_int64)]
        l.zb_scene_create.argtypes = [vp, i, vp, i, vp, i, vp, i, vp, i, i, vp, vp, i, f, i, vp, vp, C.POINTER(vp)]
        l.zb_scene_destroy.argtypes = [vp]
        l.zb_scene_destroy.restype = None
        l.zb_scene_update_robin_coefficients.argtypes = [vp, vp, vp]
        l.zb_scene_update_robin_coefficients_dev.argtypes = [vp, vp, vp]
        l.zb_scene_stats.argtypes = [vp, i, C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        l.zb_context_set_bvh_builder.argtypes = [vp, i]
        l.zb_scene_update_vertices.argtypes = [vp, i, vp, i]
        l.zb_scene_update_vertices_dev.argtypes = [vp, i, vp, i]
        l.zb_scene_build_profile.argtypes = [vp, i, C.POINTER(C.c_float), C.POINTER(C.c_int32)]
        l.zb_scene_validate.argtypes = [vp, i, vp, i, vp, i, C.POINTER(C.c_int64)]
        l.zb_scene_download_records.argtypes = [vp, i, vp, vp, vp]
        l.zb_closest_point.argtypes = [vp, i, i] + [vp]*6
        l.zb_intersect_ray.argtypes = [vp, i, i] + [vp]*7
        l.zb_occluded.argtypes = [vp, i, i] + [vp]*3
        l.zb_count_ray_hits.argtypes = [vp, i, i] + [vp]*4
        l.zb_inside_domain.argtypes = [vp, i, vp, vp]
        l.zb_sample_sphere.argtypes = [vp, i, i] + [vp]*7
        l.zb_star_radius.argtypes = [vp, i, vp, vp, vp, f, vp, vp]
        l.zb_solve.argtypes = [vp, i, vp, vp, vp, vp, vp]
        l.zb_solve_dev.argtypes = [vp, i, vp, vp, vp, vp, vp]
        try:
            l.zb_solve_channels.argtypes = [vp, i, i, vp, vp, vp, vp, vp]
            l.zb_solve_channels_dev.argtypes = [vp, i, i, vp, vp, vp, vp, vp]
        except AttributeError:
            if "ZOMBIE_B200_LIB" not in os.environ:   # only an older A/B variant library may lack them
                raise
        l.zb_host_bvh_check.argtypes = [i, vp, i, vp, i, i, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                        C.POINTER(C.c_int64)]
        l.zb_bvc_splat.argtypes = [vp, i, f, vp, f, f, f, f, f, vp, C.POINTER(C.c_int64)]
        l.zb_bvc_splat_dev.argtypes = [vp, i, f, vp, f, f, f, f, f, vp, C.POINTER(C.c_int64)]
        l.zb_context_last_splat_ms.argtypes = [vp, C.POINTER(C.c_float)]
        l.zb_rws_solve.argtypes = [vp, vp, vp, vp, f, f, f, vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        l.zb_grid_create.argtypes = [vp, vp, C.POINTER(C.c_int32 * 3), C.POINTER(vp)]
        l.zb_grid_data.argtypes = [vp]
        l.zb_grid_data.restype = vp
        l.zb_grid_destroy.argtypes = [vp]
        l.zb_grid_destroy.restype = None
        l.zb_pcg32_kat.argtypes = [vp, C.c_uint64, C.c_uint64, i, vp]
        l.zb_pcg32_demo.argtypes = [vp, C.c_uint64, C.c_uint64, i, vp, vp, vp, vp]
        l.zb_pcg32_floats.argtypes = [vp, C.c_uint64, C.c_uint64, i, vp]
        _lib = l
    return _lib


def _check(rc, what):
    if rc != 0:
        raise ZombieError("%s failed (%d): %s" % (what, rc, lib().zb_last_error().decode()))


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def device_count():
    return lib().zb_device_count()


def host_bvh_check(dim, positions, indices, reflecting=False):
    """zb_host_bvh_check: builds the wide BVH on the host and validates it (no device needed)."""
    pos = _f32(positions)
    idx = np.ascontiguousarray(indices, np.int32)
    n, l, d, v = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int64()
    _check(lib().zb_host_bvh_check(dim, abi.ptr(pos), len(pos), abi.ptr(idx), len(idx), int(reflecting),
                                   C.byref(n), C.byref(l), C.byref(d), C.byref(v)), "zb_host_bvh_check")
    return {"nodes": n.value, "leaves": l.value, "max_depth": d.value, "violations": v.value}


class Grid:
    """zb_grid: a HostGrid's samples resident on a context's device."""

    def __init__(self, ctx, host_grid):
        self._h = C.c_void_p()
        nz, ny, nx = host_grid.values.shape
        shape = (C.c_int32 * 3)(nx, ny, nz)
        _check(lib().zb_grid_create(ctx._h, abi.ptr(host_grid.values), C.byref(shape), C.byref(self._h)), "zb_grid_create")
        self.data = lib().zb_grid_data(self._h)

    def close(self):
        if self._h:
            lib().zb_grid_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def device_pde(ctx, pde):
    """the descriptor zb_solve takes: grids attached with PdeDesc.set_grids are uploaded once per context (cached on the
    HostGrid) and their views point at device memory"""
    grids = getattr(pde, "host_grids", None)
    if not grids:
        return pde
    d = abi.PdeDesc()
    C.memmove(C.byref(d), C.byref(pde), C.sizeof(abi.PdeDesc))
    keep = []
    for name, g in grids.items():
        cache = g.__dict__.setdefault("_device", {})
        if id(ctx) not in cache:
            cache[id(ctx)] = Grid(ctx, g)
        keep.append(cache[id(ctx)])
        setattr(d, name, g.view(cache[id(ctx)].data))
    d.__dict__["_keep"] = keep
    return d


class Context:
    """zb_context: one per process and device."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        self._scenes = weakref.WeakSet()   # a scene must be destroyed before its context, whatever order the GC picks
        _check(lib().zb_context_create(device, C.byref(self._h)), "zb_context_create")

    def close(self):
        if self._h:
            for s in list(self._scenes):
                s.close()
            lib().zb_context_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def sm_count(self):
        return lib().zb_context_sm_count(self._h)

    @property
    def device(self):
        return lib().zb_context_device(self._h)

    def set_bvh_builder(self, builder):
        """zb_context_set_bvh_builder: abi.BVH_BUILDER_HOST_SAH (default) or abi.BVH_BUILDER_DEVICE for the scenes created afterwards"""
        _check(lib().zb_context_set_bvh_builder(self._h, int(builder)), "zb_context_set_bvh_builder")

    def set_stream(self, cuda_stream):
        _check(lib().zb_context_set_stream(self._h, C.c_void_p(cuda_stream)), "zb_context_set_stream")

    def synchronize(self):
        _check(lib().zb_context_synchronize(self._h), "zb_context_synchronize")

    def last_solve_profile(self):
        p, w, r = C.c_float(), C.c_float(), C.c_float()
        walks, grid, launches = C.c_int64(), C.c_int32(), C.c_int64()
        _check(lib().zb_context_last_solve_profile(self._h, C.byref(p), C.byref(w), C.byref(r), C.byref(walks),
                                                   C.byref(grid), C.byref(launches)), "zb_context_last_solve_profile")
        return {"prepare_ms": p.value, "walk_ms": w.value, "reduce_ms": r.value, "walks": walks.value,
                "walk_grid": grid.value, "launches": launches.value}

    def bvc_splat(self, dim, absorption_coeff, samples, eval_pts, radius_clamp=1e-3, kernel_regularization=0.0,
                  robin_cutoff=FLT_MAX, cutoff_a=0.0, cutoff_r=0.0, device_eval_struct=None):
        """BoundaryValueCaching::splat (boundary_value_caching.h:482-502): samples = abi.HostBvcSamples, eval_pts =
        abi.HostBvcEvalPoints (updated in place). With device_eval_struct (an abi.BvcEvalPoints of device pointers)
        the evaluation points stay resident on the device. Returns the number of pairs evaluated."""
        pairs = C.c_int64(0)
        ss = samples.as_struct()
        if device_eval_struct is not None:
            _check(lib().zb_bvc_splat_dev(self._h, dim, absorption_coeff, C.byref(ss), radius_clamp, kernel_regularization,
                                          robin_cutoff, cutoff_a, cutoff_r, C.byref(device_eval_struct), C.byref(pairs)),
                   "zb_bvc_splat_dev")
        else:
            es = eval_pts.as_struct()
            _check(lib().zb_bvc_splat(self._h, dim, absorption_coeff, C.byref(ss), radius_clamp, kernel_regularization,
                                      robin_cutoff, cutoff_a, cutoff_r, C.byref(es), C.byref(pairs)), "zb_bvc_splat")
        return pairs.value

    def last_splat_ms(self):
        ms = C.c_float()
        _check(lib().zb_context_last_splat_ms(self._h, C.byref(ms)), "zb_context_last_splat_ms")
        return ms.value

    def pcg32(self, initstate, initseq, n):
        out = np.zeros(n, np.uint32)
        _check(lib().zb_pcg32_kat(self._h, initstate, initseq, n, abi.ptr(out)), "zb_pcg32_kat")
        return out


    def pcg32_demo(self, initstate, initseq, rounds):
        u = np.zeros((rounds, 6), np.uint32)
        coins = np.zeros((rounds, 65), np.uint8)
        rolls = np.zeros((rounds, 33), np.uint8)
        cards = np.zeros((rounds, 52), np.uint8)
        _check(lib().zb_pcg32_demo(self._h, initstate, initseq, rounds, abi.ptr(u), abi.ptr(coins), abi.ptr(rolls),
                                   abi.ptr(cards)), "zb_pcg32_demo")
        return u, coins, rolls, cards

    def pcg32_floats(self, initstate, initseq, n):
        out = np.zeros(n, np.float32)
        _check(lib().zb_pcg32_floats(self._h, initstate, initseq, n, abi.ptr(out)), "zb_pcg32_floats")
        return out


class Scene:
    """zb_scene: the Fcpw{Dirichlet,Neumann,Robin}BoundaryHandler + GeometricQueries of the reference
    (zombie/utils/fcpw_geometric_queries.h:84-150, 1063-1291) on the device."""

    def __init__(self, ctx, dim, absorbing=None, reflecting=None, reflecting_kind=abi.REFLECTING_NONE,
                 min_robin=None, max_robin=None, solve_double_sided=False, silhouette_precision=1e-3,
                 watertight=True, domain_min=None, domain_max=None, builder=None):
        """builder: abi.BVH_BUILDER_HOST_SAH / abi.BVH_BUILDER_DEVICE for this scene (None: the context's setting)"""
        self.ctx = ctx
        self.dim = dim
        if dim not in (2, 3):
            raise ZombieError("zb_scene_create: dim must be 2 or 3")
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
        self._h = C.c_void_p()
        if builder is None and os.environ.get("ZOMBIE_B200_BVH_BUILDER"):  # A/B runs of unmodified scripts: "device" / "host"
            builder = abi.BVH_BUILDER_DEVICE if os.environ["ZOMBIE_B200_BVH_BUILDER"] == "device" else abi.BVH_BUILDER_HOST_SAH
        if builder is not None:
            ctx.set_bvh_builder(builder)
        try:
            _check(lib().zb_scene_create(ctx._h, dim, abi.ptr(ap), len(ap), abi.ptr(ai), len(ai),
                                         abi.ptr(rp), len(rp), abi.ptr(ri), len(ri),
                                         reflecting_kind, abi.ptr(mn), abi.ptr(mx),
                                         int(solve_double_sided), silhouette_precision, int(watertight),
                                         abi.ptr(self.domain_min), abi.ptr(self.domain_max), C.byref(self._h)),
                   "zb_scene_create")
        finally:
            if builder is not None:
                ctx.set_bvh_builder(abi.BVH_BUILDER_HOST_SAH)
        ctx._scenes.add(self)

    def close(self):
        if self._h:
            if self.ctx._h:   # (a closing context destroys its scenes first)
                lib().zb_scene_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def stats(self, which):
        b, n, l, d = C.c_int64(), C.c_int32(), C.c_int32(), C.c_int32()
        _check(lib().zb_scene_stats(self._h, which, C.byref(b), C.byref(n), C.byref(l), C.byref(d)), "zb_scene_stats")
        return {"bytes": b.value, "nodes": n.value, "leaves": l.value, "max_depth": d.value}

    def update_vertices(self, which, positions):
        """zb_scene_update_vertices: moving geometry — new positions of all vertices of one boundary; the hierarchy is rebuilt
        on the device"""
        pos = _f32(positions)
        _check(lib().zb_scene_update_vertices(self._h, which, abi.ptr(pos), len(pos)), "zb_scene_update_vertices")

    def update_vertices_dev(self, which, positions_ptr, n_vertices):
        """the same with a DEVICE pointer to n_vertices*dim floats"""
        _check(lib().zb_scene_update_vertices_dev(self._h, which, C.c_void_p(positions_ptr), int(n_vertices)),
               "zb_scene_update_vertices_dev")

    def build_profile(self, which):
        ms, b = C.c_float(), C.c_int32()
        _check(lib().zb_scene_build_profile(self._h, which, C.byref(ms), C.byref(b)), "zb_scene_build_profile")
        return {"ms": ms.value, "builder": b.value}

    def validate(self, which, positions, indices):
        """zb_scene_validate (test hook): violations of the hierarchy as it sits on the device, against the caller's mesh"""
        pos = _f32(positions)
        idx = np.ascontiguousarray(indices, np.int32)
        v = C.c_int64(-1)
        _check(lib().zb_scene_validate(self._h, which, abi.ptr(pos), len(pos), abi.ptr(idx), len(idx), C.byref(v)), "zb_scene_validate")
        return v.value

    def download_records(self, which, n_prims, adjacency=False):
        """zb_scene_download_records (test hook): leaf-order primitive index, vertex records and adjacency records"""
        prim = np.zeros(n_prims, np.int32)
        recs = np.zeros((n_prims, 12 if self.dim == 3 else 4), np.float32)
        adj = np.zeros((n_prims, 12), np.float32) if adjacency else None
        _check(lib().zb_scene_download_records(self._h, which, abi.ptr(prim), abi.ptr(recs), abi.ptr(adj) if adjacency else None),
               "zb_scene_download_records")
        return prim, recs, adj

    def update_robin_coefficients(self, min_robin, max_robin):
        mn, mx = _f32(min_robin), _f32(max_robin)
        _check(lib().zb_scene_update_robin_coefficients(self._h, abi.ptr(mn), abi.ptr(mx)), "zb_scene_update_robin_coefficients")

    def update_robin_coefficients_dev(self, min_ptr, max_ptr):
        """the same with DEVICE pointers (e.g. tensor.data_ptr() of float32 tensors, one value per reflecting primitive);
        asynchronous on the context's stream"""
        _check(lib().zb_scene_update_robin_coefficients_dev(self._h, C.c_void_p(min_ptr), C.c_void_p(max_ptr)),
               "zb_scene_update_robin_coefficients_dev")

    def closest_point(self, which, x, max_radius=None, want_normal=True):
        x = _f32(x)
        n = len(x)
        prim = np.zeros(n, np.int32)
        dist = np.zeros(n, np.float32)
        pt = np.zeros((n, self.dim), np.float32)
        nrm = np.zeros((n, self.dim), np.float32) if want_normal else None
        _check(lib().zb_closest_point(self._h, which, n, abi.ptr(x), abi.ptr(_f32(max_radius)),
                                      abi.ptr(prim), abi.ptr(dist), abi.ptr(pt), abi.ptr(nrm)), "zb_closest_point")
        return prim, dist, pt, nrm

    def intersect_ray(self, which, origin, direction, tmax=None):
        o, d = _f32(origin), _f32(direction)
        n = len(o)
        prim = np.zeros(n, np.int32)
        dist = np.zeros(n, np.float32)
        pt = np.zeros((n, self.dim), np.float32)
        nrm = np.zeros((n, self.dim), np.float32)
        _check(lib().zb_intersect_ray(self._h, which, n, abi.ptr(o), abi.ptr(d), abi.ptr(_f32(tmax)),
                                      abi.ptr(prim), abi.ptr(dist), abi.ptr(pt), abi.ptr(nrm)), "zb_intersect_ray")
        return prim, dist, pt, nrm

    def occluded(self, which, a, b):
        a, b = _f32(a), _f32(b)
        out = np.zeros(len(a), np.uint8)
        _check(lib().zb_occluded(self._h, which, len(a), abi.ptr(a), abi.ptr(b), abi.ptr(out)), "zb_occluded")
        return out

    def count_ray_hits(self, which, origin, direction, tmax=None):
        """number of distinct hits of each ray (FCPW's all-hits intersection, duplicates merged)"""
        o, d = _f32(origin), _f32(direction)
        t = None if tmax is None else _f32(tmax)
        out = np.zeros(len(o), np.int32)
        _check(lib().zb_count_ray_hits(self._h, which, len(o), abi.ptr(o), abi.ptr(d), abi.ptr(t) if t is not None else None, abi.ptr(out)),
               "zb_count_ray_hits")
        return out

    def inside_domain(self, x):
        """GeometricQueries::insideDomain for a batch of points (bool array)"""
        x = _f32(x)
        out = np.zeros(len(x), np.uint8)
        _check(lib().zb_inside_domain(self._h, len(x), abi.ptr(x), abi.ptr(out)), "zb_inside_domain")
        return out.astype(bool)

    def sample_sphere(self, x, radius, rand_nums, which=abi.REFLECTING):
        x, radius, u = _f32(x), _f32(radius), _f32(rand_nums)
        n = len(x)
        prim = np.zeros(n, np.int32)
        pt = np.zeros((n, self.dim), np.float32)
        nrm = np.zeros((n, self.dim), np.float32)
        pdf = np.zeros(n, np.float32)
        _check(lib().zb_sample_sphere(self._h, which, n, abi.ptr(x), abi.ptr(radius), abi.ptr(u),
                                      abi.ptr(prim), abi.ptr(pt), abi.ptr(nrm), abi.ptr(pdf)), "zb_sample_sphere")
        return prim, pt, nrm, pdf

    def star_radius(self, x, min_radius, max_radius, precision=1e-3, flip=None):
        x, mn, mx = _f32(x), _f32(min_radius), _f32(max_radius)
        out = np.zeros(len(x), np.float32)
        fl = None if flip is None else np.ascontiguousarray(flip, np.uint8)
        _check(lib().zb_star_radius(self._h, len(x), abi.ptr(x), abi.ptr(mn), abi.ptr(mx), precision,
                                    abi.ptr(fl), abi.ptr(out)), "zb_star_radius")
        return out

    def solve(self, estimator, pde, settings, pts, stats):
        """WalkOnSpheres::solve / WalkOnStars::solve (vector overloads) with host arrays; pts and stats
        are updated in place like the reference's samplePts / statistics. Returns the walk steps taken."""
        steps = C.c_int64(0)
        ps, ss = pts.as_struct(), stats.as_struct()
        pde = device_pde(self.ctx, pde)
        _check(lib().zb_solve(self._h, estimator, C.byref(pde), C.byref(settings), C.byref(ps), C.byref(ss),
                              C.byref(steps)), "zb_solve")
        return steps.value

    def solve_channels(self, estimator, pdes, settings, pts, stats_list):
        """zb_solve_channels: the reference's vector-valued solve (T = Array<float, C>): one set of walks, one PdeDesc and
        one abi.HostStatistics per channel. Returns the walk steps taken."""
        if len(pdes) != len(stats_list) or not pdes:
            raise ZombieError("solve_channels: one statistics object per PDE descriptor")
        steps = C.c_int64(0)
        ps = pts.as_struct()
        dev = [device_pde(self.ctx, p) for p in pdes]
        pde_arr = (abi.PdeDesc * len(dev))()
        for k, d in enumerate(dev):
            C.memmove(C.byref(pde_arr, k*C.sizeof(abi.PdeDesc)), C.byref(d), C.sizeof(abi.PdeDesc))
        st_arr = (abi.Statistics * len(stats_list))(*[s.as_struct() for s in stats_list])
        _check(lib().zb_solve_channels(self._h, estimator, len(dev), pde_arr, C.byref(settings), C.byref(ps), st_arr, C.byref(steps)),
               "zb_solve_channels")
        return steps.value

    def solve_channels_dev(self, estimator, pdes, settings, pts_struct, stats_structs, steps_dev_ptr=None):
        """zb_solve_channels_dev: the same with every array resident on the device (one abi.Statistics of device pointers per channel)"""
        dev = [device_pde(self.ctx, p) for p in pdes]
        pde_arr = (abi.PdeDesc * len(dev))()
        for k, d in enumerate(dev):
            C.memmove(C.byref(pde_arr, k*C.sizeof(abi.PdeDesc)), C.byref(d), C.sizeof(abi.PdeDesc))
        st_arr = (abi.Statistics * len(stats_structs))(*stats_structs)
        _check(lib().zb_solve_channels_dev(self._h, estimator, len(dev), pde_arr, C.byref(settings), C.byref(pts_struct), st_arr,
                                           C.c_void_p(steps_dev_ptr)), "zb_solve_channels_dev")

    def rws_solve(self, pde, settings, samples, eval_pts, normal_offset, radius_clamp=0.0, kernel_regularization=0.0):
        """rws::ReverseWalkOnStarsSolver::solve (reverse_walk_solver.h:287-330): samples = abi.HostRwsSamples (rngState
        advances), eval_pts = abi.HostRwsEvalPoints (totals accumulate). Returns (walk steps, splat records)."""
        steps, records = C.c_int64(0), C.c_int64(0)
        ss, es = samples.as_struct(), eval_pts.as_struct()
        pde = device_pde(self.ctx, pde)
        _check(lib().zb_rws_solve(self._h, C.byref(pde), C.byref(settings), C.byref(ss), normal_offset, radius_clamp,
                                  kernel_regularization, C.byref(es), C.byref(steps), C.byref(records)), "zb_rws_solve")
        return steps.value, records.value

    def solve_dev(self, estimator, pde, settings, pts_struct, stats_struct, steps_dev_ptr=None):
        """Same with every array already resident on the device (abi.SamplePoints / abi.Statistics of
        device pointers); results are ordered on the context's stream (the call enqueues walk rounds until the walks have ended)."""
        pde = device_pde(self.ctx, pde)
        _check(lib().zb_solve_dev(self._h, estimator, C.byref(pde), C.byref(settings), C.byref(pts_struct),
                                  C.byref(stats_struct), C.c_void_p(steps_dev_ptr)), "zb_solve_dev")
