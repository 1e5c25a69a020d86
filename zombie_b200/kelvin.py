"""Host pre/post-processing of exterior problems: zombie::KelvinTransform
(zombie/variance_reduction/kelvin_transform.h:109-121, 244-279, 351-388), vectorised with numpy. The PDE
wrapper itself (transformPde, :282-348) runs on the device: set PdeDesc.kelvin = 1 and kelvinOrigin, describe
the PDE of the EXTERIOR domain, hand zb_scene_create the inverted mesh and zb_solve the inverted points.
Everything is float32 like the reference."""
import numpy as np

F = np.float32


def transform_points(x, origin):
    """KelvinTransform::transformPoint (:109-113): y = (x - o)/|x - o|^2"""
    s = np.asarray(x, F) - np.asarray(origin, F)
    return (s/(s*s).sum(axis=1, dtype=F, keepdims=True)).astype(F)


def transform_normals(x, n, origin):
    """KelvinTransform::transformNormal (:116-121)"""
    s = np.asarray(x, F) - np.asarray(origin, F)
    n = np.asarray(n, F)
    return (n - F(2)*(n*s).sum(axis=1, keepdims=True)*s/(s*s).sum(axis=1, keepdims=True)).astype(F)


def _closest_point_on_triangle(pa, pb, pc, x):
    """computeClosestPointOnTriangle (zombie/core/geometry_helpers.h:94-161; Ericson 5.1.5), vectorised"""
    ab, ac, ax = pb - pa, pc - pa, x - pa
    d1, d2 = (ab*ax).sum(1), (ac*ax).sum(1)
    bx = x - pb
    d3, d4 = (ab*bx).sum(1), (ac*bx).sum(1)
    cx = x - pc
    d5, d6 = (ab*cx).sum(1), (ac*cx).sum(1)
    vc = d1*d4 - d3*d2
    vb = d5*d2 - d1*d6
    va = d3*d6 - d5*d4
    with np.errstate(divide="ignore", invalid="ignore"):
        v_ab = d1/(d1 - d3)
        w_ac = d2/(d2 - d6)
        w_bc = (d4 - d3)/((d4 - d3) + (d5 - d6))
        denom = 1.0/(va + vb + vc)
    pt = pa + ab*(vb*denom)[:, None] + ac*(vc*denom)[:, None]
    # regions in reverse order of precedence so that earlier tests of the reference win
    m = (va <= 0) & ((d4 - d3) >= 0) & ((d5 - d6) >= 0)
    pt = np.where(m[:, None], pb + (pc - pb)*w_bc[:, None], pt)
    m = (vb <= 0) & (d2 >= 0) & (d6 <= 0)
    pt = np.where(m[:, None], pa + ac*w_ac[:, None], pt)
    m = (vc <= 0) & (d1 >= 0) & (d3 <= 0)
    pt = np.where(m[:, None], pa + ab*v_ab[:, None], pt)
    m = (d6 >= 0) & (d5 <= d6)
    pt = np.where(m[:, None], pc, pt)
    m = (d3 >= 0) & (d4 <= d3)
    pt = np.where(m[:, None], pb, pt)
    m = (d1 <= 0) & (d2 <= 0)
    pt = np.where(m[:, None], pa, pt)
    return pt.astype(F)


def _closest_point_on_segment(pa, pb, x):
    u, v = pb - pa, x - pa
    c1, c2 = (u*v).sum(1), (u*u).sum(1)
    with np.errstate(divide="ignore", invalid="ignore"):
        t = c1/c2
    pt = pa + u*t[:, None]
    pt = np.where((c2 <= c1)[:, None], pb, pt)
    pt = np.where((c1 <= 0)[:, None], pa, pt)
    return pt.astype(F)


def compute_robin_coefficients(transformed_points, indices, min_coeff, max_coeff):
    """KelvinTransform::computeRobinCoefficients (:244-279): per-primitive bounds on the Robin coefficient of the
    inverted problem, mu/r^2 (+ n.y/r^2 in 3D), from the closest and farthest points of the primitive to the
    origin of the inverted domain."""
    p = np.asarray(transformed_points, F)
    idx = np.asarray(indices, np.int64)
    dim = p.shape[1]
    mn, mx = np.asarray(min_coeff, F), np.asarray(max_coeff, F)
    zero = np.zeros((len(idx), dim), F)
    verts = [p[idx[:, k]] for k in range(dim)]
    if dim == 3:
        closest = _closest_point_on_triangle(verts[0], verts[1], verts[2], zero)
    else:
        closest = _closest_point_on_segment(verts[0], verts[1], zero)
    d_closest = np.sqrt((closest*closest).sum(1, dtype=F))
    d2 = np.stack([(v*v).sum(1, dtype=F) for v in verts], axis=1)
    if dim == 3:  # computeFarthestPointOnTriangle: first maximal vertex in the order a, b, c with >= (geometry_helpers.h:163-180)
        pick = np.where((d2[:, 0] >= d2[:, 1]) & (d2[:, 0] >= d2[:, 2]), 0, np.where((d2[:, 1] >= d2[:, 0]) & (d2[:, 1] >= d2[:, 2]), 1, 2))
    else:         # computeFarthestPointOnLineSegment: a only if strictly farther (geometry_helpers.h:51-63)
        pick = np.where(d2[:, 0] > d2[:, 1], 0, 1)
    farthest = np.stack(verts, axis=1)[np.arange(len(idx)), pick]
    d_farthest = np.sqrt(d2[np.arange(len(idx)), pick])
    min_shift = np.zeros(len(idx), F)
    max_shift = np.zeros(len(idx), F)
    if dim == 3:
        n = np.cross(verts[1] - verts[0], verts[2] - verts[0]).astype(F)
        n = n/np.sqrt((n*n).sum(1, dtype=F, keepdims=True))
        dc, df = (n*closest).sum(1, dtype=F), (n*farthest).sum(1, dtype=F)
        min_shift, max_shift = np.minimum(dc, df), np.maximum(dc, df)
    min_value = (mn + min_shift)/(d_farthest*d_farthest)
    max_value = (mx + max_shift)/(d_closest*d_closest)
    eps = np.finfo(F).eps
    out_min = np.where(min_value*max_value > eps, np.minimum(np.abs(min_value), np.abs(max_value)), F(0))
    out_max = np.maximum(np.abs(min_value), np.abs(max_value))
    return out_min.astype(F), out_max.astype(F)


def transform_solution_estimate(V, y):
    """KelvinTransform::transformSolutionEstimate (:351-356): u(x) = |y| V(y) in 3D, V(y) in 2D"""
    y = np.asarray(y, F)
    V = np.asarray(V, F)
    return (np.sqrt((y*y).sum(1, dtype=F))*V).astype(F) if y.shape[1] == 3 else V


def transform_gradient_estimate(V, dV, y):
    """KelvinTransform::transformGradientEstimate (:359-388)"""
    y = np.asarray(y, F)
    V = np.asarray(V, F)
    dV = np.asarray(dV, F)
    dim = y.shape[1]
    r2 = (y*y).sum(1, dtype=F)
    r_pow = np.sqrt(r2) if dim == 3 else np.ones_like(r2)
    dU = (r_pow*r2)[:, None]*dV
    if dim == 3:
        dU = dU + (r_pow*V)[:, None]*y
    R = np.eye(dim, dtype=F)[None] - F(2)*y[:, :, None]*y[:, None, :]/r2[:, None, None]
    return np.einsum("nij,nj->ni", R, dU).astype(F)
