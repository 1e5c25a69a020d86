"""Seeded inputs shared by the CPU (oracle) and GPU (parity) tests and by tests/golden/make_golden.py."""
import math

import numpy as np

from zombie_b200 import abi, procedural as P


def mesh3d(subdiv=3):
    """bumpy icosphere split into Dirichlet (z <= 0) and reflecting (z > 0) parts"""
    pos, tri = P.bumpy_sphere(subdiv)
    absorbing, reflecting = P.partition_mesh(pos, tri, lambda c: c[:, 2] > 0)
    return (pos, tri), absorbing, reflecting


def mesh2d(n_segments=512):
    pos, seg = P.star_polygon(n_segments)
    absorbing, reflecting = P.partition_mesh(pos, seg, lambda c: c[:, 1] > 0)
    return (pos, seg), absorbing, reflecting


def query_points(n, dim, radius, seed):
    return P.points_in_ball(n, radius, seed=seed, dim=dim)


def rays(n, dim, radius, seed):
    rng = np.random.RandomState(seed)
    o = P.points_in_ball(n, radius, seed=seed + 1, dim=dim)
    d = rng.normal(size=(n, dim))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    tmax = rng.uniform(0.05, 2.0, n)
    return o, np.ascontiguousarray(d, np.float32), np.ascontiguousarray(tmax, np.float32)


def scene_kwargs(dim, kind, absorbing, reflecting, mu=1.0, double_sided=False):
    kw = dict(absorbing=absorbing, reflecting=reflecting, reflecting_kind=kind, solve_double_sided=double_sided)
    if kind == abi.REFLECTING_ROBIN:
        c = np.full(len(reflecting[1]), mu, np.float32)
        kw.update(min_robin=c, max_robin=c)
    return kw


def quadratic_pde(mu=0.0, sigma=0.0, dim=3):
    """u = x^2 - y^2/2 (+ z^2/4 in 3D) + 0.3; lap u = 1.5 in 3D, 1 in 2D"""
    pz = 0.25 if dim == 3 else 0.0
    return abi.PdeDesc.manufactured(abi.SOL_QUADRATIC, [1.0, -0.5, pz, 0.3], absorptionCoeff=sigma, robinCoeff=mu)


def quadratic_exact(x):
    u = x[:, 0]**2 - 0.5*x[:, 1]**2 + 0.3
    if x.shape[1] == 3:
        u = u + 0.25*x[:, 2]**2
    return u


def exp_pde(mu=1.0, sigma=1.0):
    k = np.array([1.0, 1.0, 1.0])/math.sqrt(3.0)*math.sqrt(sigma)
    return abi.PdeDesc.manufactured(abi.SOL_EXP_DIR, list(k) + [1.0], absorptionCoeff=sigma, robinCoeff=mu), k.astype(np.float32)


def z_scores(a, b):
    """per-point z of the difference of two independent estimates given as HostStatistics"""
    va = a.solution_variance()/np.maximum(1, a.nSolutionEstimates)
    vb = b.solution_variance()/np.maximum(1, b.nSolutionEstimates)
    return (a.solutionMean - b.solutionMean)/np.sqrt(np.maximum(va + vb, 1e-12))


def _distance_to_primitives(p, x):
    """distance from x to every segment (p: [n, 2, 2]) / triangle (p: [n, 3, 3]), float64, vectorised"""
    def seg(a, b):
        ab = b - a
        t = np.clip(np.einsum("ij,ij->i", x[None, :] - a, ab)/np.maximum(np.einsum("ij,ij->i", ab, ab), 1e-300), 0.0, 1.0)
        return np.linalg.norm(a + ab*t[:, None] - x[None, :], axis=1)
    if p.shape[2] == 2:
        return seg(p[:, 0], p[:, 1])
    a, b, c = p[:, 0], p[:, 1], p[:, 2]
    n = np.cross(b - a, c - a)
    nn = np.maximum(np.einsum("ij,ij->i", n, n), 1e-300)
    q = x[None, :] - n*(np.einsum("ij,ij->i", x[None, :] - a, n)/nn)[:, None]      # projection onto the plane
    inside = np.ones(len(p), bool)
    for u, v in ((a, b), (b, c), (c, a)):
        inside &= np.einsum("ij,ij->i", np.cross(v - u, q - u), n) >= 0.0
    d_plane = np.abs(np.einsum("ij,ij->i", x[None, :] - a, n))/np.sqrt(nn)
    d_edges = np.minimum(np.minimum(seg(a, b), seg(b, c)), seg(c, a))
    return np.where(inside, d_plane, d_edges)


def coplanar_primitive_within(positions, indices, x, lo, hi, rel):
    """True when x lies in the supporting plane (3D) / line (2D) of a primitive whose distance from x is in [lo, hi]:
    plane distance below `rel` times that distance. float64 throughout."""
    p = positions[indices]                      # (n, dim, dim)
    dim = p.shape[2]
    if dim == 3:
        n = np.cross(p[:, 1] - p[:, 0], p[:, 2] - p[:, 0])
    else:
        e = p[:, 1] - p[:, 0]
        n = np.stack([e[:, 1], -e[:, 0]], axis=1)
    n /= np.maximum(np.linalg.norm(n, axis=1, keepdims=True), 1e-300)
    h = np.abs(np.einsum("ij,ij->i", n, x[None, :] - p[:, 0]))
    d = _distance_to_primitives(p, x)
    near = (d >= lo) & (d <= hi)
    return bool(np.any(near & (h <= rel*np.maximum(d, 1e-300))))
