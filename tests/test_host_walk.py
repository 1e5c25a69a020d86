"""CPU cross-check of the DEVICE walk code: zombie_b200/csrc/{walk,queries,greens,pde,rng}.cuh are host+device, and
tools/host_solve.cu runs them one walk at a time on the CPU (test tool only — the library has no CPU path). Here its
estimates, walk lengths and first-sphere radii are compared with the reference's TBB solver (oracle/_ref) and the
exact solution, so estimator regressions are caught without a GPU. Same gates as tests/test_gpu_solve.py.

(This test found a real defect: FCPW's 2D SilhouetteVertex test and its 3D SilhouetteEdge test use opposite view-
direction conventions relative to Zombie's reflectance primitives; applying the 3D sign in 2D made walks along 2D
Neumann boundaries 35 % too short.)"""
import os
import struct
import subprocess

import numpy as np
import pytest

from tests import cases
from tests.conftest import ROOT
from zombie_b200 import abi


@pytest.fixture(scope="module")
def host_solve(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("host") / "host_solve")
    cmd = ["/usr/local/cuda/bin/nvcc", "-std=c++17", "-O2", "-DZB_HOST_STATS", "-Xcompiler", "-ffp-contract=off", "-o", exe,
           os.path.join(ROOT, "tools", "host_solve.cu"), os.path.join(ROOT, "zombie_b200", "csrc", "wide_bvh.cpp")]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout[-3000:]
    return exe


def run_host(exe, d, dim, kind, absorbing, reflecting, x, n_walks, family, params, sigma, mu, rr, split, seed,
             types=None, quantities=None, normals=None, directions=None, double_sided=False, kelvin_origin=None,
             robin_bounds=None, sides=None, channels=None):
    hdr = struct.pack("<4i16f4fQ", dim, kind, n_walks, family, *(list(params) + [0.0]*(16 - len(params))), sigma, mu, rr, split, seed)
    open(os.path.join(d, "hdr.bin"), "wb").write(hdr)
    empty = (np.zeros((0, dim), np.float32), np.zeros((0, dim), np.int32))
    for name, (pos, idx) in (("a", absorbing), ("r", reflecting if reflecting is not None else empty)):
        np.ascontiguousarray(pos, np.float32).tofile(os.path.join(d, name + "pos.bin"))
        np.ascontiguousarray(idx, np.int32).tofile(os.path.join(d, name + "idx.bin"))
    np.ascontiguousarray(x, np.float32).tofile(os.path.join(d, "pts.bin"))
    for name, arr, dt in (("type", types, np.int32), ("quantity", quantities, np.int32), ("normal", normals, np.float32),
                          ("dir", directions, np.float32)):
        path = os.path.join(d, name + ".bin")
        if arr is not None:
            np.ascontiguousarray(arr, dt).tofile(path)
        elif os.path.exists(path):
            os.remove(path)
    np.array([1 if double_sided else 0], np.int32).tofile(os.path.join(d, "double_sided.bin"))
    np.array([0.0]*4 if kelvin_origin is None else [1.0] + [float(v) for v in kelvin_origin], np.float32).tofile(os.path.join(d, "kelvin.bin"))
    for name, arr in (("rmin", None if robin_bounds is None else robin_bounds[0]), ("rmax", None if robin_bounds is None else robin_bounds[1])):
        path = os.path.join(d, name + ".bin")
        if arr is not None:
            np.ascontiguousarray(arr, np.float32).tofile(path)
        elif os.path.exists(path):
            os.remove(path)
    # optional: constant Dirichlet data of the two sides of a double-sided boundary; parameter sets of extra value channels
    for name, arr in (("sides", sides), ("channels", None if channels is None else np.asarray(channels, np.float32).reshape(-1, 16))):
        path = os.path.join(d, name + ".bin")
        if arr is not None:
            np.ascontiguousarray(arr, np.float32).tofile(path)
        elif os.path.exists(path):
            os.remove(path)
    r = subprocess.run([exe, d], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0 and "status 0" in r.stdout, r.stdout
    out = np.fromfile(os.path.join(d, "out.bin"), np.float32).reshape(-1, 6)
    st = abi.HostStatistics(len(x), dim)
    st.solutionMean, st.solutionM2 = out[:, 0].copy(), out[:, 1].copy()
    st.nSolutionEstimates = out[:, 2].astype(np.int32)
    st.totalWalkLength = out[:, 3].astype(np.int32)
    st.totalDerivativeContribution = out[:, 4].copy()
    st.firstSphereRadius = out[:, 5].copy()
    return st


CASES = {
    "wost2d_neumann": dict(dim=2, kind=abi.REFLECTING_NEUMANN, family=abi.SOL_QUADRATIC, params=[1.0, -0.5, 0.0, 0.3], sigma=0.0, mu=0.0),
    "wost3d_neumann": dict(dim=3, kind=abi.REFLECTING_NEUMANN, family=abi.SOL_QUADRATIC, params=[1.0, -0.5, 0.25, 0.3], sigma=0.0, mu=0.0),
    "wost3d_robin_yukawa": dict(dim=3, kind=abi.REFLECTING_ROBIN, family=abi.SOL_EXP_DIR, params=[3**-0.5]*3 + [1.0], sigma=1.0, mu=1.0),
    "wost2d_robin": dict(dim=2, kind=abi.REFLECTING_ROBIN, family=abi.SOL_QUADRATIC, params=[1.0, -0.5, 0.0, 0.3], sigma=0.0, mu=0.7),
    # screened Poisson in 2D: the Bessel-function ball Green's function (distributions.h:678-833)
    # Walk on Spheres (walk_on_spheres.h:229-288): Dirichlet boundary only; 2D Laplace with C1's cubic, 3D Poisson
    "wos2d": dict(dim=2, kind=abi.REFLECTING_NONE, family=abi.SOL_CUBIC_2D, params=[0.0]*4, sigma=0.0, mu=0.0),
    "wos3d_poisson": dict(dim=3, kind=abi.REFLECTING_NONE, family=abi.SOL_QUADRATIC, params=[1.0, -0.5, 0.25, 0.3], sigma=0.0, mu=0.0),
    # the weight window of the demo configs (engine/wost.json: Russian roulette below 0.99, splitting above 1.5;
    # walk_on_stars.h:310-343) on a problem whose throughput falls (screening) and rises (reflectance <= 1 only): roulette
    "wost3d_robin_yukawa_window": dict(dim=3, kind=abi.REFLECTING_ROBIN, family=abi.SOL_EXP_DIR, params=[3**-0.5]*3 + [1.0], sigma=1.0,
                                       mu=1.0, rr=0.99, split=1.5),
    "wost2d_neumann_yukawa": dict(dim=2, kind=abi.REFLECTING_NEUMANN, family=abi.SOL_EXP_DIR, params=[0.6, 0.8, 0.0, 1.0], sigma=1.0, mu=0.0),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_device_walk_code_on_host_matches_reference(ref, host_solve, tmp_path, name):
    c = CASES[name]
    dim = c["dim"]
    whole, absorbing, reflecting = cases.mesh3d(3) if dim == 3 else cases.mesh2d(512)
    wos = c["kind"] == abi.REFLECTING_NONE
    if wos:
        absorbing, reflecting = whole, None
    x = cases.query_points(160, dim, 0.6 if dim == 3 else 0.7, seed=77)
    n_walks = 192
    rr, split = c.get("rr", 0.0), c.get("split", abi.FLT_MAX)
    st = run_host(host_solve, str(tmp_path), dim, c["kind"], absorbing, reflecting, x, n_walks, c["family"], c["params"],
                  c["sigma"], c["mu"], rr, split, 0xBEEF)
    rs = ref.RefScene(dim, absorbing=absorbing) if wos else ref.RefScene(dim, **cases.scene_kwargs(dim, c["kind"], absorbing, reflecting, mu=c["mu"]))
    _, da, _, _ = rs.closest_point(abi.ABSORBING, x)
    dr = None if wos else rs.closest_point(abi.REFLECTING, x)[1]
    pts = abi.HostSamplePoints(x, da, dr, nWalks=n_walks, estimationQuantity=abi.SOLUTION, seed=0xF00D)
    rst = abi.HostStatistics(len(x), dim)
    pde = abi.PdeDesc.manufactured(c["family"], c["params"], absorptionCoeff=c["sigma"], robinCoeff=c["mu"])
    rsteps, _ = rs.solve(abi.WALK_ON_SPHERES if wos else abi.WALK_ON_STARS, pde,
                         abi.WalkSettings.default(russianRouletteThreshold=rr, splittingThreshold=split), pts, rst)
    assert np.array_equal(st.nSolutionEstimates, rst.nSolutionEstimates)
    assert abs(st.totalWalkLength.sum()/float(rsteps) - 1.0) < 0.04      # same walk-length distribution
    z = cases.z_scores(st, rst)
    assert np.mean(np.abs(z) < 3.0) >= 0.98
    assert abs(z.mean()) < 0.3 and 0.75 < z.std() < 1.3, (z.mean(), z.std())


def test_device_walk_code_on_host_double_sided(ref, host_solve, tmp_path):
    """WalkSettings::solveDoubleSided (walk_on_stars.h:357-376, 419-436; fcpw_geometric_queries.h:1029-1038): an open
    reflecting plate inside a Dirichlet sphere, walks passing on both of its sides; no edge is ignored as convex, the
    plate's rim consists of edges without an adjacent face (always silhouettes), normals flip with the side the walk is
    on. u = x + 0.2 is harmonic with zero normal derivative on the plate z = 0.05."""
    whole, _, _ = cases.mesh3d(3)
    g = np.linspace(-0.4, 0.4, 13, dtype=np.float32)
    gx, gy = np.meshgrid(g, g, indexing="ij")
    ppos = np.stack([gx.ravel(), gy.ravel(), np.full(gx.size, 0.05, np.float32)], axis=1).astype(np.float32)
    n = len(g)
    quads = [(i*n + j, (i + 1)*n + j, (i + 1)*n + j + 1, i*n + j + 1) for i in range(n - 1) for j in range(n - 1)]
    ptri = np.array([t for a, b, c, d in quads for t in ((a, b, c), (a, c, d))], np.int32)
    plate = (ppos, ptri)
    x = cases.query_points(112, 3, 0.55, seed=91)
    x = x[np.abs(x[:, 2] - 0.05) > 0.02]                       # not inside the plate's epsilon shell
    n_walks = 128
    family, params = abi.SOL_LINEAR, [1.0, 0.0, 0.0, 0.2]
    st = run_host(host_solve, str(tmp_path), 3, abi.REFLECTING_NEUMANN, whole, plate, x, n_walks, family, params, 0.0, 0.0, 0.0,
                  abi.FLT_MAX, 0xBEEF, double_sided=True)
    rs = ref.RefScene(3, **cases.scene_kwargs(3, abi.REFLECTING_NEUMANN, whole, plate, double_sided=True))
    _, da, _, _ = rs.closest_point(abi.ABSORBING, x)
    _, dr, _, _ = rs.closest_point(abi.REFLECTING, x)
    pts = abi.HostSamplePoints(x, da, dr, nWalks=n_walks, estimationQuantity=abi.SOLUTION, seed=0xF00D)
    rst = abi.HostStatistics(len(x), 3)
    pde = abi.PdeDesc.manufactured(family, params, absorptionCoeff=0.0, robinCoeff=0.0)
    rsteps, _ = rs.solve(abi.WALK_ON_STARS, pde, abi.WalkSettings.default(solveDoubleSided=True), pts, rst)
    assert np.array_equal(st.nSolutionEstimates, rst.nSolutionEstimates)
    assert abs(st.totalWalkLength.sum()/float(rsteps) - 1.0) < 0.04
    z = cases.z_scores(st, rst)
    assert np.mean(np.abs(z) < 3.0) >= 0.98
    assert abs(z.mean()) < 0.3 and 0.75 < z.std() < 1.3, (z.mean(), z.std())
    # and both agree with the exact solution
    exact = x[:, 0] + 0.2
    se = np.sqrt(st.solutionM2/np.maximum(st.nSolutionEstimates - 1, 1)/np.maximum(st.nSolutionEstimates, 1))
    assert np.mean(np.abs(st.solutionMean - exact) < 4.0*se + 2e-3) >= 0.97


@pytest.mark.parametrize("mixed", [False, True])
def test_device_walk_code_on_host_kelvin_exterior(ref, host_solve, tmp_path, mixed):
    """BASELINE configs[4]: exterior Laplace problem on the Kelvin-inverted domain, the PDE wrapped on
    the device like KelvinTransform::transformPde (kelvin_transform.h:282-348; pde.cuh). Dirichlet everywhere (Walk on
    Spheres), and Neumann data on z > 0 that the inversion turns into Robin conditions with the per-triangle coefficient
    bounds of computeRobinCoefficients (Walk on Stars, coefficients of either sign)."""
    from zombie_b200 import kelvin, workloads
    w = workloads.config5(subdivisions=3, n_points=144, n_walks=128, mixed=mixed)
    kw = w.scene_kwargs()
    bounds = (kw["min_robin"], kw["max_robin"]) if mixed else None
    st = run_host(host_solve, str(tmp_path), 3, abi.REFLECTING_ROBIN if mixed else abi.REFLECTING_NONE, kw["absorbing"],
                  kw.get("reflecting"), w.points, w.n_walks, abi.SOL_INV_DIST, [float(v) for v in w.pde.params[:4]], 0.0, 0.0, 0.0,
                  abi.FLT_MAX, 0xBEEF, kelvin_origin=w.kelvin_origin, robin_bounds=bounds)
    rs = ref.RefScene(3, **kw)
    _, da, _, _ = rs.closest_point(abi.ABSORBING, w.points)
    dr = rs.closest_point(abi.REFLECTING, w.points)[1] if mixed else None
    pts = abi.HostSamplePoints(w.points, da, dr, nWalks=w.n_walks, estimationQuantity=abi.SOLUTION, seed=0xF00D)
    rst = abi.HostStatistics(pts.n, 3)
    rsteps, _ = rs.solve(w.estimator, w.pde, w.settings, pts, rst)
    assert np.array_equal(st.nSolutionEstimates, rst.nSolutionEstimates)
    assert abs(st.totalWalkLength.sum()/float(rsteps) - 1.0) < 0.05
    z = cases.z_scores(st, rst)
    assert np.mean(np.abs(z) < 3.0) >= 0.97
    assert abs(z.mean()) < 0.35 and 0.7 < z.std() < 1.35, (z.mean(), z.std())
    if not mixed:
        u = kelvin.transform_solution_estimate(st.solutionMean, w.points)
        se = np.linalg.norm(w.points, axis=1)*np.sqrt(st.solutionM2/np.maximum(st.nSolutionEstimates - 1, 1)/np.maximum(st.nSolutionEstimates, 1))
        assert np.mean(np.abs(u - w.exterior_exact(w.exterior_points)) < 4.0*se + 2e-3) > 0.95


def test_cooperative_traversal_on_emulated_warps(tmp_path):
    """zombie_b200/csrc/coop.cuh (the warp-cooperative closest-point / star-radius traversal the walk kernel runs) on
    a CPU: tools/coop_check.cu emulates each warp with 32 host threads (tools/warp_emu.h) and compares every query
    with the one-thread-per-query traversal: bit-identical distances, star radii equal to float rounding, no
    work-list overflow — for a Robin and a Neumann boundary."""
    exe = str(tmp_path / "coop_check")
    cmd = ["/usr/local/cuda/bin/nvcc", "-std=c++17", "-O2", "-DZB_EMULATE_WARP", "-Xcompiler", "-ffp-contract=off", "-Xcompiler",
           "-pthread", "-o", exe, os.path.join(ROOT, "tools", "coop_check.cu"), os.path.join(ROOT, "zombie_b200", "csrc", "wide_bvh.cpp")]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout[-3000:]
    # Robin coefficients from "the plane-distance bound rejects most triangles" (0.1) to "the reflectance bound sets
    # nearly every radius" (100): the filter / full rounds of the pooled phase and the padded pruning bounds must not
    # change a single result
    for mu in ("1.0", "0", "0.1", "100"):
        r = subprocess.run([exe, "5", "1024", mu], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        assert "distance mismatches 0" in r.stdout and "cooperative status 0" in r.stdout, r.stdout
        assert "star: mismatches 0" in r.stdout, r.stdout
        worst = float(r.stdout.split("worst rel ")[1].split(")")[0])
        assert worst < 2e-5, r.stdout   # squared radius; fused vs reference-ordered primitive arithmetic
    # the child-per-lane form of the node rounds (a build option, measured slower on the device) returns the same results
    exe2 = str(tmp_path / "coop_check_by_child")
    r = subprocess.run(cmd[:-4] + ["-DZB_STAR_CHILD_PARALLEL=1", "-DZB_CLOSEST_CHILD_PARALLEL=1", "-o", exe2] + cmd[-2:], stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout[-3000:]
    r = subprocess.run([exe2, "5", "1024", "1.0"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert "distance mismatches 0" in r.stdout and "cooperative status 0" in r.stdout and "star: mismatches 0" in r.stdout, r.stdout


def test_cooperative_traversal_2d_on_emulated_warps(tmp_path):
    """The DIM = 2 instantiation of coop.cuh (line segments, silhouette vertices; configs C1 / C2) on emulated warps
    (tools/coop_check2d.cu): identical distances and star radii to the one-thread traversal on a star-shaped polygon,
    Robin and Neumann, coarse and fine."""
    exe = str(tmp_path / "coop_check2d")
    cmd = ["/usr/local/cuda/bin/nvcc", "-std=c++17", "-O2", "-DZB_EMULATE_WARP", "-Xcompiler", "-ffp-contract=off", "-Xcompiler",
           "-pthread", "-o", exe, os.path.join(ROOT, "tools", "coop_check2d.cu"), os.path.join(ROOT, "zombie_b200", "csrc", "wide_bvh.cpp")]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout[-3000:]
    for args in (("4096", "1024", "1"), ("4096", "1024", "0"), ("512", "512", "5")):
        r = subprocess.run([exe, *args], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        assert r.returncode == 0 and "distance mismatches 0" in r.stdout and "star: mismatches 0" in r.stdout, r.stdout
        assert "cooperative status 0" in r.stdout, r.stdout


@pytest.fixture(scope="module")
def host_query(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("hostq") / "host_query")
    cmd = ["/usr/local/cuda/bin/nvcc", "-std=c++17", "-O2", "-DZB_HOST_STATS", "-Xcompiler", "-ffp-contract=off", "-o", exe,
           os.path.join(ROOT, "tools", "host_query.cu"), os.path.join(ROOT, "zombie_b200", "csrc", "wide_bvh.cpp")]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout[-3000:]
    return exe


@pytest.mark.parametrize("dim,kind,mu", [(2, abi.REFLECTING_NEUMANN, 0.0), (2, abi.REFLECTING_ROBIN, 2.0),
                                         (3, abi.REFLECTING_NEUMANN, 0.0), (3, abi.REFLECTING_ROBIN, 2.0)])
def test_device_star_radius_code_on_host_matches_reference(ref, host_query, tmp_path, dim, kind, mu):
    """GeometricQueries::computeStarRadiusForReflectingBoundary (fcpw_geometric_queries.h:1250-1291): the device
    traversals compiled for the host against the reference's BVH on interior points and on points ON the boundary
    (where walks continue after a reflection). Where the reference's BVH and its own brute-force Baseline disagree by
    more than FCPW's 1e-3 tolerance (a few queries in 10^5), the device code must side with the Baseline."""
    whole, absorbing, reflecting = cases.mesh3d(4) if dim == 3 else cases.mesh2d(2048)
    rp, ri = reflecting
    rng = np.random.RandomState(3)
    n_b = 6000
    k = rng.randint(0, len(ri), n_b)
    w = rng.dirichlet([1.0]*dim, n_b).astype(np.float32)
    on_boundary = sum(rp[ri[k, j]]*w[:, j:j + 1] for j in range(dim)).astype(np.float32)
    x = np.concatenate([cases.query_points(30000, dim, 0.99, seed=41), on_boundary])
    check_star_radius_against_reference(ref, host_query, tmp_path, dim, kind, mu, absorbing, reflecting, x)


def check_star_radius_against_reference(ref, host_query, tmp_path, dim, kind, mu, absorbing, reflecting, x, min_shrunk=0.05):
    rp, ri = reflecting
    kw = cases.scene_kwargs(dim, kind, absorbing, reflecting, mu=mu)
    rs = ref.RefScene(dim, **kw)
    _, da, _, _ = rs.closest_point(abi.ABSORBING, x)
    _, dr, _, _ = rs.closest_point(abi.REFLECTING, x)
    d = str(tmp_path)
    x.tofile(os.path.join(d, "pts.bin"))
    da.astype(np.float32).tofile(os.path.join(d, "maxr.bin"))
    np.ascontiguousarray(rp, np.float32).tofile(os.path.join(d, "rpos.bin"))
    np.ascontiguousarray(ri, np.int32).tofile(os.path.join(d, "ridx.bin"))
    r = subprocess.run([host_query, d, str(dim), str(kind), str(mu)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    out = np.fromfile(os.path.join(d, "query_out.bin"), np.float32).reshape(-1, 3)
    assert np.array_equal(out[:, 0], dr)                         # closest-point distance: bit-identical to FCPW
    # primitive index: the reference's Baseline aggregate on every query (ties go to the largest input index, queries.cuh)
    rb = ref.RefScene(dim, use_baseline=True, **kw)
    bprim, bd, _, _ = rb.closest_point(abi.REFLECTING, x)
    rb.close()
    gprim = np.fromfile(os.path.join(d, "query_prim.bin"), np.int32)
    assert np.array_equal(bd, dr) and np.array_equal(gprim, bprim)
    rr = rs.star_radius(x, np.full(len(x), 1e-3, np.float32), da)
    for col, name in ((1, "trigonometric node test"), (2, "algebraic node test + fast primitive test")):
        g = out[:, col]
        bad = np.abs(rr - g) > 1e-3
        if bad.any():   # every such row must be one where the reference's BVH deviates from its own Baseline
            rb = ref.RefScene(dim, use_baseline=True, **kw)
            base = rb.star_radius(x[bad], np.full(int(bad.sum()), 1e-3, np.float32), da[bad])
            assert (np.abs(base - g[bad]) <= 1e-5).all(), (name, rr[bad], g[bad], base)
        assert np.mean(np.abs(rr - g)[~bad] <= 1e-5*np.maximum(rr[~bad], 1e-3) + 1e-7) > 0.99, name
    assert (rr < 0.999*da).mean() > min_shrunk


FCPW_INPUT = "/root/reference/deps/fcpw/tests/input"


def load_obj(path, dim):
    """vertices and the triangles ('f', fan-triangulated) or segments ('l' polylines) of a Wavefront file"""
    pos, idx = [], []
    for line in open(path):
        t = line.split()
        if not t:
            continue
        if t[0] == "v":
            pos.append([float(v) for v in t[1:1 + dim]])
        elif t[0] == "f" and dim == 3:
            v = [int(a.split("/")[0]) - 1 for a in t[1:]]
            idx += [[v[0], v[i], v[i + 1]] for i in range(1, len(v) - 1)]
        elif t[0] == "l" and dim == 2:
            v = [int(a) - 1 for a in t[1:]]
            idx += [[v[i], v[i + 1]] for i in range(len(v) - 1)]
    return np.array(pos, np.float32), np.array(idx, np.int32)


@pytest.mark.skipif(not os.path.isdir(FCPW_INPUT), reason="FCPW's test meshes live in the reference checkout")
@pytest.mark.parametrize("name,dim,kind,mu", [("bunny.obj", 3, abi.REFLECTING_NEUMANN, 0.0), ("bunny.obj", 3, abi.REFLECTING_ROBIN, 3.0),
                                              ("walker.obj", 2, abi.REFLECTING_NEUMANN, 0.0), ("walker.obj", 2, abi.REFLECTING_ROBIN, 3.0)])
def test_device_query_code_on_fcpw_test_meshes(ref, host_query, tmp_path, name, dim, kind, mu):
    """The meshes FCPW's own differential tests run on (deps/fcpw/tests/aggregate_tests.cpp, tests/input/): scanned,
    irregular triangles and an open 2D polyline instead of this repo's procedural shapes. The mesh is cut by a
    coordinate plane into an absorbing and a reflecting part (open boundaries along the cut: edges without an adjacent
    face are always silhouettes); closest-point distances must be bit-identical to FCPW, star radii equal under the
    same rule as above. CPU only: the reference checkout does not travel to the GPU box."""
    pos, idx = load_obj(os.path.join(FCPW_INPUT, name), dim)
    lo, hi = pos.min(0), pos.max(0)
    pos = ((pos - 0.5*(lo + hi))/(0.5*(hi - lo).max())).astype(np.float32)   # into [-1, 1]^dim
    from zombie_b200 import procedural
    absorbing, reflecting = procedural.partition_mesh(pos, idx, lambda c: c[:, dim - 1] > 0.0)  # compacted vertex arrays
    rp, ri = reflecting
    rng = np.random.RandomState(5)
    n_b = 3000
    k = rng.randint(0, len(ri), n_b)
    w = rng.dirichlet([1.0]*dim, n_b).astype(np.float32)
    on_boundary = sum(rp[ri[k, j]]*w[:, j:j + 1] for j in range(dim)).astype(np.float32)
    box = rng.uniform(-1.0, 1.0, (12000, dim)).astype(np.float32)*(pos.max(0) - pos.min(0))*0.5 + 0.5*(pos.max(0) + pos.min(0))
    x = np.concatenate([box.astype(np.float32), on_boundary])
    check_star_radius_against_reference(ref, host_query, tmp_path, dim, kind, mu, absorbing, reflecting, x, min_shrunk=0.02)


def test_device_walk_code_on_host_at_cache_points(ref, host_solve, tmp_path):
    """The sample points Boundary Value Caching solves at (boundary_value_caching.h:567-626): points ON the reflecting
    boundary (walks start with a hemispherical step, walk_on_stars.h:410-416) and points displaced inward from the
    absorbing boundary whose normal derivative is estimated as a directional derivative (walk_on_stars.h:665-848)."""
    from zombie_b200 import bvc
    whole, absorbing, reflecting = cases.mesh3d(3)
    sa, sr = bvc.UniformBoundarySampler(*absorbing, seed=3), bvc.UniformBoundarySampler(*reflecting, seed=4)
    sa.initialize(0.03, False)
    sr.initialize(0.0, False)
    a = sa.generate_samples(96, abi.ON_ABSORBING_BOUNDARY, 0.03)
    r = sr.generate_samples(96, abi.ON_REFLECTING_BOUNDARY, 0.0)
    x = np.concatenate([a.pt, r.pt])
    normals = np.concatenate([a.normal, r.normal])
    types = np.concatenate([a.type, r.type])
    quantities = np.concatenate([np.full(96, abi.SOLUTION_AND_GRADIENT, np.int32), np.full(96, abi.SOLUTION, np.int32)])
    n_walks = 256
    st = run_host(host_solve, str(tmp_path), 3, abi.REFLECTING_NEUMANN, absorbing, reflecting, x, n_walks, abi.SOL_QUADRATIC,
                  [1.0, -0.5, 0.25, 0.3], 0.0, 0.0, 0.0, abi.FLT_MAX, 0xBEEF, types=types, quantities=quantities,
                  normals=normals, directions=normals)
    rs = ref.RefScene(3, **cases.scene_kwargs(3, abi.REFLECTING_NEUMANN, absorbing, reflecting))
    _, da, _, _ = rs.closest_point(abi.ABSORBING, x)
    _, dr, _, _ = rs.closest_point(abi.REFLECTING, x)
    pts = abi.HostSamplePoints(x, da, dr, nWalks=n_walks, estimationQuantity=quantities, type=types, normal=normals, seed=0xF00D)
    pts.directionForDerivative = np.ascontiguousarray(normals)
    rst = abi.HostStatistics(len(x), 3)
    rsteps, _ = rs.solve(abi.WALK_ON_STARS, cases.quadratic_pde(), abi.WalkSettings.default(), pts, rst)
    assert np.array_equal(st.nSolutionEstimates, rst.nSolutionEstimates)
    assert abs(st.totalWalkLength.sum()/float(rsteps) - 1.0) < 0.04
    assert np.allclose(st.firstSphereRadius, pts.firstSphereRadius, rtol=1e-5, atol=1e-7)
    z = cases.z_scores(st, rst)
    assert np.mean(np.abs(z) < 3.0) >= 0.97 and abs(z.mean()) < 0.3 and 0.75 < z.std() < 1.3, (z.mean(), z.std())
    # getEstimatedDerivative (common.h:341-345) at the displaced absorbing points: same mean over the batch
    dh = st.totalDerivativeContribution[:96]/np.maximum(1, st.nSolutionEstimates[:96])
    dref = rst.totalDerivativeContribution[:96]/np.maximum(1, rst.nSolutionEstimates[:96])
    exact = (np.stack([2.0*x[:96, 0], -x[:96, 1], 0.5*x[:96, 2]], 1)*normals[:96]).sum(1)
    se = np.sqrt(np.var(dh - exact)/96 + np.var(dref - exact)/96)
    assert abs(np.mean(dh - exact) - np.mean(dref - exact)) < 4.0*se + 0.02
    assert abs(np.mean(dh - exact)) < 4.0*np.std(dh - exact)/np.sqrt(96) + 0.02


@pytest.mark.parametrize("dim", [2, 3])
def test_device_ray_and_sampling_code_on_host_matches_reference(ref, tmp_path, dim):
    """Aggregate::intersect(Ray) closest hit / occlusion (bvh.inl:656-771) and the stochastic sphere-sampling descent
    (bvh.inl:1141-1288) as the kernels run them, compiled for the host (tools/host_rays.cu): hits against the compiled
    reference under the gates of the GPU test (primitive identical or an exact tie in t, t and point to 1e-6, normal to
    1e-5, occlusion flags identical), sampling as an importance sampler (E[1/pdf] per query within 4 standard errors of
    the reference's own sampler; the pdf depends on the tree, so nothing finer is comparable)."""
    from tests.test_gpu_queries import check_rays
    exe = str(tmp_path / "host_rays")
    cmd = ["/usr/local/cuda/bin/nvcc", "-std=c++17", "-O2", "-Xcompiler", "-ffp-contract=off", "-o", exe,
           os.path.join(ROOT, "tools", "host_rays.cu"), os.path.join(ROOT, "zombie_b200", "csrc", "wide_bvh.cpp")]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout[-3000:]
    whole, absorbing, reflecting = cases.mesh3d(4) if dim == 3 else cases.mesh2d(2048)
    pos, idx = reflecting
    nq = 20000
    o, d, tmax = cases.rays(nq, dim, 0.9, seed=5)
    n = 100000
    sx = np.repeat(cases.query_points(50, dim, 0.8, seed=51), n//50, axis=0)
    rad = np.full(n, 0.6, np.float32)
    u = np.random.RandomState(5).uniform(size=(n, dim)).astype(np.float32)
    dd = str(tmp_path)
    for name, a, t in (("pos", pos, np.float32), ("idx", idx, np.int32), ("o", o, np.float32), ("d", d, np.float32),
                       ("tmax", tmax, np.float32), ("sx", sx, np.float32), ("srad", rad, np.float32), ("su", u, np.float32)):
        np.ascontiguousarray(a, t).tofile(os.path.join(dd, name + ".bin"))
    r = subprocess.run([exe, dd, str(dim)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    rays = np.fromfile(os.path.join(dd, "rays_out.bin"), np.float32).reshape(nq, 3 + 2*dim)
    smp = np.fromfile(os.path.join(dd, "samples_out.bin"), np.float32).reshape(n, 2 + 2*dim)

    rs = ref.RefScene(dim, **cases.scene_kwargs(dim, abi.REFLECTING_NEUMANN, absorbing, reflecting))
    ref_hits = rs.intersect_ray(abi.REFLECTING, o, d, tmax)
    assert (ref_hits[0] >= 0).mean() > 0.1 and (ref_hits[0] < 0).mean() > 0.1
    check_rays(ref_hits, (rays[:, 0].astype(np.int32), rays[:, 1], rays[:, 2:2 + dim], rays[:, 2 + dim:2 + 2*dim]))
    assert np.array_equal(rs.occluded(abi.REFLECTING, o, o + d*tmax[:, None]).astype(bool), rays[:, 2 + 2*dim] > 0.5)

    # all-hits counts (FCPW's recordAllHits intersection with duplicates merged; the parity of the axis-aligned count is
    # what GeometricQueries::insideDomain reads)
    hits = np.fromfile(os.path.join(dd, "hits_out.bin"), np.int32).reshape(nq, 2)
    ex = np.zeros_like(d)
    ex[:, 0] = 1.0
    assert np.array_equal(hits[:, 0], rs.count_ray_hits(abi.REFLECTING, o, d, tmax))
    assert np.array_equal(hits[:, 1], rs.count_ray_hits(abi.REFLECTING, o, ex))
    assert (hits[:, 1] == 0).mean() > 0.02 and (hits[:, 1] == 1).mean() > 0.02 and (dim == 3 or (hits >= 2).mean() > 0.005)

    # axis-parallel rays THROUGH mesh vertices (what insideDomain casts from a point level with a vertex): up to six
    # triangles report the vertex at distances a few ulps apart and merge into one hit only as a chain, and the ray lies in a
    # face plane of the leaf boxes (0*inf in a product-form slab test)
    vo = pos[np.random.RandomState(3).randint(0, len(pos), 2000)].astype(np.float32)
    vo[:, 0] -= np.float32(0.37)
    for name, a in (("o", vo), ("d", d[:2000]), ("tmax", tmax[:2000])):
        np.ascontiguousarray(a, np.float32).tofile(os.path.join(dd, name + ".bin"))
    r = subprocess.run([exe, dd, str(dim)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    vhits = np.fromfile(os.path.join(dd, "hits_out.bin"), np.int32).reshape(2000, 2)
    assert np.array_equal(vhits[:, 1], rs.count_ray_hits(abi.REFLECTING, vo, ex[:2000]))
    assert (vhits[:, 1] >= 1).all()

    rprim, rpt, rn, rpdf = rs.sample_sphere(sx, rad, u)
    gprim, gpdf, gpt, gn = smp[:, 0].astype(np.int32), smp[:, 1], smp[:, 2:2 + dim], smp[:, 2 + dim:]
    f = gprim >= 0
    assert f.mean() > 0.2 and (gpdf[f] > 0).all()
    bad = 0
    for j in range(50):
        sl = slice(j*(n//50), (j + 1)*(n//50))
        wr = np.where(rprim[sl] >= 0, 1.0/np.maximum(rpdf[sl], 1e-30), 0.0)
        wg = np.where(gprim[sl] >= 0, 1.0/np.maximum(gpdf[sl], 1e-30), 0.0)
        se = np.sqrt(wr.var()/len(wr) + wg.var()/len(wg))
        bad += abs(wr.mean() - wg.mean()) > 4.0*se + 1e-6
    assert bad <= 2
    a = pos[idx[gprim[f], 0]]
    assert np.abs(((gpt[f] - a)*gn[f]).sum(axis=1)).max() < 1e-5   # the sample lies on the primitive it names
    rs.close()


def test_value_channels_on_host_equal_scalar_runs(host_solve, tmp_path):
    """zb_solve_channels' device code (walk.cuh: extra*Term, reduceChannel; the CH instantiations) on the CPU: two extra channels
    on the walks of a Robin / screened-Poisson gradient solve equal scalar runs of their PDEs on the same generators"""
    c = CASES["wost3d_robin_yukawa"]
    whole, absorbing, reflecting = cases.mesh3d(2)
    x = cases.query_points(24, 3, 0.6, seed=4)
    k = 3**-0.5
    extra = [[k, -k, k, 0.5] + [0.0]*12, [0.0, 0.6, 0.8, -2.0] + [0.0]*12]
    q = np.full(len(x), abi.SOLUTION_AND_GRADIENT, np.int32)
    d = str(tmp_path)
    run_host(host_solve, d, 3, c["kind"], absorbing, reflecting, x, 16, c["family"], c["params"], c["sigma"], c["mu"], 0.0, abi.FLT_MAX, 77,
             quantities=q, channels=extra)
    multi = [np.fromfile(os.path.join(d, "out_c%d.bin" % j), np.float32).reshape(-1, 6) for j in range(2)]
    for j, params in enumerate(extra):
        st = run_host(host_solve, d, 3, c["kind"], absorbing, reflecting, x, 16, c["family"], params[:4], c["sigma"], c["mu"], 0.0, abi.FLT_MAX, 77,
                      quantities=q)
        assert np.array_equal(multi[j][:, 2].astype(np.int32), st.nSolutionEstimates)
        assert np.allclose(multi[j][:, 0], st.solutionMean, rtol=0, atol=2e-5*max(1.0, float(np.abs(st.solutionMean).max())))
        assert np.allclose(multi[j][:, 1], st.solutionM2, rtol=1e-3, atol=1e-5)
    assert not np.allclose(multi[0][:, 0], multi[1][:, 0])


def test_two_sided_dirichlet_data_on_host_match_reference(host_solve, tmp_path, ref):
    """the side flag of double-sided solves (walk.cuh: onNormalAlignedSide -> pde.dirichlet(x, aligned)) on the CPU against the
    compiled reference: constant data 1 / 3 on the two sides of a closed curve"""
    from tests.test_double_sided_dirichlet import constant_pde, problem, scene_kwargs, solve
    pos, seg, x, inside = problem(n=300)
    r = solve(ref.RefScene(2, **scene_kwargs(pos, seg)), constant_pde(), x, n_walks=16)
    st = run_host(host_solve, str(tmp_path), 2, abi.REFLECTING_NONE, (pos, seg), None, x, 16, abi.SOL_CUBIC_2D, [0.0]*4, 0.0, 0.0, 0.0, abi.FLT_MAX, 5,
                  double_sided=True, sides=[1.0, 3.0])
    ok = (st.nSolutionEstimates > 0) & (r.nSolutionEstimates > 0)
    ins, out = st.solutionMean[inside & ok], st.solutionMean[~inside & ok]
    assert (ins == 1.0).mean() > 0.8 and (ins < 1.5).all() and (out == 3.0).mean() > 0.8 and (out > 2.5).all()
    assert abs(out.mean() - r.solutionMean[~inside & ok].mean()) < 0.03 and abs(ins.mean() - r.solutionMean[inside & ok].mean()) < 0.03


def test_value_channels_on_host_walk_on_spheres_2d(host_solve, tmp_path):
    """the same for Walk on Spheres in 2D (sampled source, no reflecting boundary): channels with other quadratic solutions"""
    (pos, seg), _, _ = cases.mesh2d(256)
    x = cases.query_points(20, 2, 0.5, seed=8)
    extra = [[0.25, 2.0, 0.0, -1.0] + [0.0]*12]
    q = np.full(len(x), abi.SOLUTION_AND_GRADIENT, np.int32)
    d = str(tmp_path)
    run_host(host_solve, d, 2, abi.REFLECTING_NONE, (pos, seg), None, x, 24, abi.SOL_QUADRATIC, [1.0, -0.5, 0.0, 0.3], 0.0, 0.0, 0.0, abi.FLT_MAX, 9,
             quantities=q, channels=extra)
    multi = np.fromfile(os.path.join(d, "out_c0.bin"), np.float32).reshape(-1, 6)
    st = run_host(host_solve, d, 2, abi.REFLECTING_NONE, (pos, seg), None, x, 24, abi.SOL_QUADRATIC, extra[0][:4], 0.0, 0.0, 0.0, abi.FLT_MAX, 9, quantities=q)
    assert np.array_equal(multi[:, 2].astype(np.int32), st.nSolutionEstimates)
    assert np.allclose(multi[:, 0], st.solutionMean, rtol=0, atol=2e-5*max(1.0, float(np.abs(st.solutionMean).max())))
    exact = 0.25*x[:, 0]**2 + 2.0*x[:, 1]**2 - 1.0
    se = np.sqrt(multi[:, 1]/np.maximum(1.0, multi[:, 2] - 1.0)/np.maximum(1.0, multi[:, 2]))
    assert np.mean(np.abs(multi[:, 0] - exact) < 4.0*se + 5e-3) > 0.9
