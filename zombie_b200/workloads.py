"""The synthetic workloads of BASELINE.json's configs, shared by bench.py, the parity
tests and the reference arm so that both sides solve the identical problem. Everything is procedural
and seed-fixed; nothing is read from disk.
"""
import math

import numpy as np

from . import abi, procedural as P


class Workload:
    """Boundary meshes + PDE + walk settings + evaluation points of one configuration."""

    def __init__(self, name, dim, absorbing, reflecting, reflecting_kind, min_robin, max_robin, pde, settings,
                 estimator, points, n_walks, quantity, exact, exact_gradient=None):
        self.name = name
        self.dim = dim
        self.absorbing = absorbing
        self.reflecting = reflecting
        self.reflecting_kind = reflecting_kind
        self.min_robin = min_robin
        self.max_robin = max_robin
        self.pde = pde
        self.settings = settings
        self.estimator = estimator
        self.points = points
        self.n_walks = n_walks
        self.quantity = quantity
        self.exact = exact                    # callable x -> u(x)
        self.exact_gradient = exact_gradient  # callable x -> grad u(x) or None

    def n_triangles(self, which):
        mesh = self.absorbing if which == abi.ABSORBING else self.reflecting
        return 0 if mesh is None else len(mesh[1])

    def scene_kwargs(self):
        return dict(absorbing=self.absorbing, reflecting=self.reflecting, reflecting_kind=self.reflecting_kind,
                    min_robin=self.min_robin, max_robin=self.max_robin,
                    solve_double_sided=bool(self.settings.solveDoubleSided),
                    silhouette_precision=float(self.settings.silhouettePrecision))

    def algorithmic_bytes_per_step(self):
        """bytes one walk step must at least read — one root-to-leaf path of the
        reference's binary tree (depth x 2 child boxes x 24 B + 4 primitives x 36 B) per query; cone
        (20 B) and Robin range (8 B) per child box for the star-radius query; the stochastic
        sphere-sample descent reads one child box per level."""
        def depth(n):
            return max(1, math.ceil(math.log2(max(2.0, n/4.0))))
        prim_bytes = 36 if self.dim == 3 else 16
        box_bytes = 24 if self.dim == 3 else 16
        na, nr = self.n_triangles(abi.ABSORBING), self.n_triangles(abi.REFLECTING)
        total = 0.0
        if na:
            total += depth(na)*2*box_bytes + 4*prim_bytes                      # closest point (A9)
        if self.estimator == abi.WALK_ON_STARS and nr:
            extra = 20 + (8 if self.reflecting_kind == abi.REFLECTING_ROBIN else 0)
            total += depth(nr)*2*(box_bytes + extra) + 4*prim_bytes            # star radius (A14/A15)
            total += depth(nr)*2*box_bytes + 4*prim_bytes                      # ray (A11)
            total += depth(nr)*box_bytes + prim_bytes                          # sphere sample (A13)
            total += depth(nr)*2*box_bytes + 4*prim_bytes                      # occlusion ray (A12)
        return total


def config3(subdivisions=8, n_points=1 << 20, n_walks=64, seed_points=7):
    """BASELINE configs[2] (the headline): 3D screened Poisson (sigma = 1) with Robin conditions (mu = 1) on
    the z > 0 half of a bumpy icosphere and Dirichlet conditions on the rest; 20*4^subdivisions triangles
    (8 -> 1,310,720); WalkOnStars with gradient estimates (antithetic + control variates) at n_points points
    uniform in the ball |x| < 0.7; manufactured solution u = exp(k.x), |k| = 1, so f = 0."""
    pos, tri = P.bumpy_sphere(subdivisions)
    absorbing, reflecting = P.partition_mesh(pos, tri, lambda c: c[:, 2] > 0)
    k = np.array([1.0, 1.0, 1.0])/math.sqrt(3.0)
    pde = abi.PdeDesc.manufactured(abi.SOL_EXP_DIR, list(k) + [1.0], absorptionCoeff=1.0, robinCoeff=1.0)
    settings = abi.WalkSettings.default(1e-3, 1e-3, 1024, False)
    mu = np.ones(len(reflecting[1]), np.float32)
    pts = P.points_in_ball(n_points, 0.7, seed=seed_points)
    kf = k.astype(np.float32)
    return Workload("C3: 3D screened Poisson, Robin z>0 / Dirichlet z<0, bumpy icosphere %d tris, WoSt + gradient, %d pts x %d walks"
                    % (len(tri), n_points, n_walks), 3, absorbing, reflecting, abi.REFLECTING_ROBIN, mu, mu, pde, settings,
                    abi.WALK_ON_STARS, pts, n_walks, abi.SOLUTION_AND_GRADIENT,
                    lambda x: np.exp(x @ kf), lambda x: np.exp(x @ kf)[:, None]*kf[None, :])


def interior_grid_2d(n, pos, margin=0.0):
    """n x n grid over the bounding box of a 2D star-shaped polygon r(theta); keeps the interior points."""
    lo, hi = pos.min(axis=0), pos.max(axis=0)
    xs = np.linspace(lo[0], hi[0], n)
    ys = np.linspace(lo[1], hi[1], n)
    g = np.stack(np.meshgrid(xs, ys, indexing="xy"), axis=-1).reshape(-1, 2)
    th = np.arctan2(g[:, 1], g[:, 0])
    r = np.linalg.norm(g, axis=1)
    inside = r < (1.0 + 0.25*np.sin(5.0*th)) - margin
    return np.ascontiguousarray(g[inside], dtype=np.float32)


def config1(n_segments=4096, grid=128, n_walks=128):
    """BASELINE configs[0]: 2D Laplace, Dirichlet, star polygon r = 1 + 0.25 sin 5 theta, u = Re (x+iy)^3,
    WalkOnSpheres at the interior points of a grid x grid lattice."""
    pos, seg = P.star_polygon(n_segments)
    pde = abi.PdeDesc.manufactured(abi.SOL_CUBIC_2D, [0.0]*4)
    settings = abi.WalkSettings.default(1e-3, 1e-3, 1024, False)
    pts = interior_grid_2d(grid, pos, margin=2e-3)
    return Workload("C1: 2D Laplace Dirichlet, star polygon %d segments, WoS, %dx%d grid x %d walks" % (n_segments, grid, grid, n_walks),
                    2, (pos, seg), None, abi.REFLECTING_NONE, None, None, pde, settings, abi.WALK_ON_SPHERES, pts, n_walks,
                    abi.SOLUTION, lambda x: x[:, 0]**3 - 3.0*x[:, 0]*x[:, 1]**2)


def config2(n_segments=4096, grid=512, n_walks=1024):
    """BASELINE configs[1]: 2D Poisson, Neumann on segments with centroid y > 0 and Dirichlet elsewhere,
    u = sin(pi x) e^(pi y) + x^2 (f = -2), WalkOnStars with the demo's weight window (rr 0.99, split 1.5)."""
    pos, seg = P.star_polygon(n_segments)
    absorbing, reflecting = P.partition_mesh(pos, seg, lambda c: c[:, 1] > 0)
    pde = abi.PdeDesc.manufactured(abi.SOL_SIN_EXP_2D, [0.0]*4)
    settings = abi.WalkSettings.default(1e-3, 1e-3, 1024, False, russianRouletteThreshold=0.99, splittingThreshold=1.5)
    pts = interior_grid_2d(grid, pos, margin=2e-3)
    pi = np.float32(math.pi)
    return Workload("C2: 2D Poisson mixed Dirichlet/Neumann, star polygon %d segments, WoSt, %dx%d grid x %d walks" % (n_segments, grid, grid, n_walks),
                    2, absorbing, reflecting, abi.REFLECTING_NEUMANN, None, None, pde, settings, abi.WALK_ON_STARS, pts, n_walks,
                    abi.SOLUTION, lambda x: np.sin(pi*x[:, 0])*np.exp(pi*x[:, 1]) + x[:, 0]**2)


def config5(subdivisions=9, n_points=1 << 20, n_walks=100, mixed=False, seed_points=11):
    """BASELINE configs[4]: 3D Laplace EXTERIOR problem outside a bumpy icosphere, solved on the Kelvin-inverted
    domain (zombie/variance_reduction/kelvin_transform.h). u = 1/|x - x0| with the pole x0 inside the body, so u is
    harmonic outside and decays like 1/r. Evaluation points lie in the shell 1.5 < |x| < 4. The PDE descriptor
    describes the EXTERIOR problem and carries kelvin = 1: the device wraps its callbacks like
    KelvinTransform::transformPde; the mesh and the points handed to the solver are the inverted ones; estimates map
    back with kelvin.transform_solution_estimate (u(x) = |y| V(y)).
    mixed = False: Dirichlet everywhere (WalkOnSpheres). mixed = True: Neumann data on faces with centroid z > 0,
    which the inversion turns into Robin conditions with coefficient n.y/|y|^2 (kelvin_transform.h:327-337), bounded
    per triangle by kelvin.compute_robin_coefficients; WalkOnStars."""
    from . import kelvin
    origin = np.zeros(3, np.float32)
    pos, tri = P.bumpy_sphere(subdivisions)
    x0 = np.array([0.2, -0.1, 0.15], np.float32)
    inv_pos = kelvin.transform_points(pos, origin)
    rng = np.random.RandomState(seed_points)
    d = rng.normal(size=(n_points, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    r = rng.uniform(1.5**3, 4.0**3, n_points)**(1.0/3.0)
    x = np.ascontiguousarray(d*r[:, None], np.float32)
    y = kelvin.transform_points(x, origin)
    pde = abi.PdeDesc.manufactured(abi.SOL_INV_DIST, list(x0) + [1.0], absorptionCoeff=0.0, robinCoeff=0.0)
    pde.kelvin = 1
    for k in range(3):
        pde.kelvinOrigin[k] = float(origin[k])
    settings = abi.WalkSettings.default(1e-3, 1e-3, 1024, False)

    def exact_inverted(yy):  # V(y) = u(x)/|y|
        xx = yy/np.sum(yy*yy, axis=1, keepdims=True) + origin
        return (1.0/np.linalg.norm(xx - x0, axis=1))/np.linalg.norm(yy, axis=1)

    if not mixed:
        w = Workload("C5: 3D Laplace exterior (Kelvin), Dirichlet, bumpy icosphere %d tris, WoS, %d pts x %d walks"
                     % (len(tri), n_points, n_walks), 3, (inv_pos, tri), None, abi.REFLECTING_NONE, None, None, pde, settings,
                     abi.WALK_ON_SPHERES, y, n_walks, abi.SOLUTION, exact_inverted)
    else:
        absorbing, reflecting = P.partition_mesh(pos, tri, lambda c: c[:, 2] > 0)
        inv_abs = (kelvin.transform_points(absorbing[0], origin), absorbing[1])
        inv_ref = (kelvin.transform_points(reflecting[0], origin), reflecting[1])
        zero = np.zeros(len(reflecting[1]), np.float32)
        mn, mx = kelvin.compute_robin_coefficients(inv_ref[0], inv_ref[1], zero, zero)
        w = Workload("C5m: 3D Laplace exterior (Kelvin), Neumann z>0 -> Robin / Dirichlet z<0, bumpy icosphere %d tris, WoSt, %d pts x %d walks"
                     % (len(tri), n_points, n_walks), 3, inv_abs, inv_ref, abi.REFLECTING_ROBIN, mn, mx, pde, settings,
                     abi.WALK_ON_STARS, y, n_walks, abi.SOLUTION, exact_inverted)
    w.exterior_points = x
    w.exterior_exact = lambda xx: 1.0/np.linalg.norm(xx - x0, axis=1)
    w.kelvin_origin = origin
    return w
