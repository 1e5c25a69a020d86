"""Procedural, seed-fixed inputs for the walk benchmark and parity tests (no files, no network).

The reference ships no 3D workload (its demo is 2D only, demo/demo.cpp:546-618), so the boundary
meshes, boundary-condition partition and evaluation points of BASELINE.json's configs are defined
here: a "bumpy" icosphere r(x) = 1 + a * sum_k sin(w_k . x + phi_k) whose
triangles are split into an absorbing (Dirichlet) and a reflecting (Neumann/Robin) part by a
centroid predicate, like zombie::partitionBoundaryMesh (utils/fcpw_geometric_queries.h:461-510).
"""
import numpy as np

_ICO_V = None


def _icosahedron():
    t = (1.0 + 5.0 ** 0.5) / 2.0
    v = np.array([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0],
                  [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t],
                  [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], dtype=np.float64)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    f = np.array([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11],
                  [1, 5, 9], [5, 11, 4], [11, 10, 2], [10, 7, 6], [7, 1, 8],
                  [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9],
                  [4, 9, 5], [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1]], dtype=np.int64)
    return v, f


def icosphere(subdivisions):
    """Unit icosphere with 20*4^s triangles, outward-facing (counter-clockwise seen from outside)."""
    v, f = _icosahedron()
    for _ in range(subdivisions):
        nv = v.shape[0]
        e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]], axis=0)
        es = np.sort(e, axis=1)
        key = es[:, 0] * nv + es[:, 1]
        uniq, inv = np.unique(key, return_inverse=True)
        a = uniq // nv
        b = uniq % nv
        mid = v[a] + v[b]
        mid /= np.linalg.norm(mid, axis=1, keepdims=True)
        v = np.concatenate([v, mid], axis=0)
        nf = f.shape[0]
        mab = nv + inv[0:nf]
        mbc = nv + inv[nf:2 * nf]
        mca = nv + inv[2 * nf:3 * nf]
        f = np.concatenate([np.stack([f[:, 0], mab, mca], 1),
                            np.stack([f[:, 1], mbc, mab], 1),
                            np.stack([f[:, 2], mca, mbc], 1),
                            np.stack([mab, mbc, mca], 1)], axis=0)
    return v, f


def bumpy_sphere(subdivisions, amplitude=0.15, seed=42, n_waves=3, max_freq=4.0):
    """Star-shaped watertight mesh r = 1 + amplitude*sum_k sin(w_k.x + phi_k) over the unit sphere.

    Returns (positions float32 [V,3], triangles int32 [T,3]); normals point out of the solid.
    """
    v, f = icosphere(subdivisions)
    rng = np.random.RandomState(seed)
    w = rng.uniform(-max_freq, max_freq, size=(n_waves, 3))
    phi = rng.uniform(0.0, 2.0 * np.pi, size=n_waves)
    r = 1.0 + amplitude * np.sin(v @ w.T + phi[None, :]).sum(axis=1) / n_waves
    p = v * r[:, None]
    return np.ascontiguousarray(p, dtype=np.float32), np.ascontiguousarray(f, dtype=np.int32)


def partition_mesh(positions, triangles, on_reflecting):
    """Splits a mesh by a predicate on primitive centroids into (absorbing, reflecting) meshes, each
    with its own compacted vertex array — the layout zombie::partitionBoundaryMesh produces."""
    c = positions[triangles].mean(axis=1)
    mask = on_reflecting(c)
    out = []
    for m in (~mask, mask):
        t = triangles[m]
        used, inv = np.unique(t.reshape(-1), return_inverse=True)
        out.append((np.ascontiguousarray(positions[used], dtype=np.float32),
                    np.ascontiguousarray(inv.reshape(-1, triangles.shape[1]), dtype=np.int32)))
    return out[0], out[1]


def flip_orientation(triangles):
    t = triangles.copy()
    t[:, [0, 1]] = t[:, [1, 0]]
    return np.ascontiguousarray(t)


def points_in_ball(n, radius, seed=7, dim=3):
    """n points uniform in the ball |x| < radius (float32)."""
    rng = np.random.RandomState(seed)
    d = rng.normal(size=(n, dim))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    r = radius * rng.uniform(size=(n, 1)) ** (1.0 / dim)
    return np.ascontiguousarray(d * r, dtype=np.float32)


def star_polygon(n_segments=4096, amplitude=0.25, lobes=5):
    """2D star-shaped polygon r(theta) = 1 + amplitude*sin(lobes*theta); counter-clockwise segments."""
    th = np.linspace(0.0, 2.0 * np.pi, n_segments, endpoint=False)
    r = 1.0 + amplitude * np.sin(lobes * th)
    p = np.stack([r * np.cos(th), r * np.sin(th)], axis=1)
    i = np.arange(n_segments)
    s = np.stack([i, (i + 1) % n_segments], axis=1)
    return np.ascontiguousarray(p, dtype=np.float32), np.ascontiguousarray(s, dtype=np.int32)


def pcg32_seed(initstate, initseq):
    """(state, inc) after pcg32::seed(initstate, initseq) (deps/pcg32/pcg32.h:53-59), vectorised."""
    mult = np.uint64(0x5851f42d4c957f2d)
    initstate = np.asarray(initstate, dtype=np.uint64)
    initseq = np.asarray(initseq, dtype=np.uint64)
    with np.errstate(over="ignore"):
        inc = (initseq << np.uint64(1)) | np.uint64(1)
        state = np.zeros_like(inc)
        state = state * mult + inc
        state = state + initstate
        state = state * mult + inc
    return state, inc


def bumpy_sphere_radius(directions, amplitude=0.15, seed=42, n_waves=3, max_freq=4.0):
    """r(d) of bumpy_sphere for unit directions d (the same waves, so the mesh interpolates this surface)"""
    rng = np.random.RandomState(seed)
    w = rng.uniform(-max_freq, max_freq, size=(n_waves, 3))
    phi = rng.uniform(0.0, 2.0 * np.pi, size=n_waves)
    d = np.asarray(directions, np.float64)
    return 1.0 + amplitude * np.sin(d @ w.T + phi[None, :]).sum(axis=1) / n_waves


def bumpy_sphere_inside(x, margin=0.0, **kw):
    """is x inside the solid r < r(x/|x|) - margin"""
    x = np.asarray(x, np.float64)
    r = np.linalg.norm(x, axis=1)
    d = x / np.maximum(r, 1e-30)[:, None]
    return r < bumpy_sphere_radius(d, **kw) - margin


def bumpy_sphere_volume(n=400000, seed=1, **kw):
    """volume of the solid: (4 pi / 3) E[r(d)^3] over uniform directions"""
    d = np.random.RandomState(seed).normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return float(4.0 * np.pi / 3.0 * np.mean(bumpy_sphere_radius(d, **kw) ** 3))
