"""The host C++ mirror of Zombie's operator API (include/zombie_b200/zombie.hpp): user code written against
PDE / GeometricQueries / Fcpw*BoundaryHandler / WalkSettings / SamplePoint / SampleStatistics /
WalkOnSpheres / WalkOnStars::solve (zombie/point_estimation/walk_on_stars.h:40-46) is compiled with g++ and,
on the GPU box, run against libzombie_b200.so; its results must equal the ones the C ABI returns for the same
seeds (the mirror only marshals AoS <-> SoA) and match the exact solution statistically."""
import ctypes
import os
import struct
import subprocess

import numpy as np
import pytest

from tests import cases
from tests.conftest import ROOT
from zombie_b200 import abi, native

DRIVER_SRC = os.path.join(ROOT, "tests", "cpp", "api_driver.cpp")


@pytest.fixture(scope="module")
def driver(tmp_path_factory):
    native.lib()  # the shared library must exist (built by __graft_entry__.build())
    exe = str(tmp_path_factory.mktemp("cpp") / "api_driver")
    libdir = os.path.join(ROOT, "zombie_b200")
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), DRIVER_SRC,
           "-L", libdir, "-lzombie_b200", "-Wl,-rpath," + libdir, "-o", exe]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    return exe


def write_problem(path, dim, estimator, kind, absorbing, reflecting, pts, n_walks, quantity, family, params,
                  sigma, mu, seed=0xC0FFEE):
    ap, ai = absorbing
    rp, ri = reflecting if reflecting is not None else (np.zeros((0, dim), np.float32), np.zeros((0, dim), np.int32))
    p16 = list(params) + [0.0]*(16 - len(params))
    hdr = struct.pack("<12i16f4f2iQ", dim, estimator, kind, len(ap), len(ai), len(rp), len(ri), len(pts), n_walks, quantity,
                      family, 0, *p16, sigma, mu, 1e-3, 1e-3, 1024, 0, seed)
    coeff = np.full(len(ri), mu, np.float32)
    with open(path, "wb") as f:
        f.write(hdr)
        for a, t in ((ap, np.float32), (ai, np.int32), (rp, np.float32), (ri, np.int32), (coeff, np.float32),
                     (coeff, np.float32), (pts, np.float32)):
            f.write(np.ascontiguousarray(a, t).tobytes())


def read_result(path, n):
    raw = open(path, "rb").read()
    steps = struct.unpack("<q", raw[:8])[0]
    extras = struct.unpack("<2f", raw[8:16])     # BVC / RWS estimates at the first point (finite when the calls ran)
    rec = np.frombuffer(raw[16:], np.float32).reshape(n, 8)
    return steps, rec, extras


def test_header_compiles_and_fails_loudly_without_a_device(driver, tmp_path):
    """CPU box: the driver builds against the header + library; with no CUDA device solve() must throw
    zombie_b200::Error (no CPU fallback). On a GPU box the same invocation succeeds."""
    _, absorbing, reflecting = cases.mesh3d(2)
    pts = cases.query_points(8, 3, 0.5, seed=3)
    prob, out = str(tmp_path / "p.bin"), str(tmp_path / "o.bin")
    write_problem(prob, 3, abi.WALK_ON_STARS, abi.REFLECTING_NEUMANN, absorbing, reflecting, pts, 4, abi.SOLUTION,
                  abi.SOL_QUADRATIC, [1.0, -0.5, 0.25, 0.3], 0.0, 0.0)
    r = subprocess.run([driver, prob, out], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    if native.lib().zb_device_count() == 0:
        assert r.returncode == 3 and "no CUDA device" in r.stderr, (r.returncode, r.stderr)
    else:
        assert r.returncode == 0, r.stderr


CASES = {
    "wos3d": dict(dim=3, estimator=abi.WALK_ON_SPHERES, kind=abi.REFLECTING_NONE, quantity=abi.SOLUTION,
                  family=abi.SOL_QUADRATIC, params=[1.0, -0.5, 0.25, 0.3], sigma=0.0, mu=0.0),
    "wost3d_robin_grad": dict(dim=3, estimator=abi.WALK_ON_STARS, kind=abi.REFLECTING_ROBIN, quantity=abi.SOLUTION_AND_GRADIENT,
                              family=abi.SOL_EXP_DIR, params=[3**-0.5, 3**-0.5, 3**-0.5, 1.0], sigma=1.0, mu=1.0),
    "wost2d_neumann": dict(dim=2, estimator=abi.WALK_ON_STARS, kind=abi.REFLECTING_NEUMANN, quantity=abi.SOLUTION,
                           family=abi.SOL_QUADRATIC, params=[1.0, -0.5, 0.0, 0.3], sigma=0.0, mu=0.0),
}


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_cpp_surface_matches_c_abi(ctx, driver, tmp_path, name):
    c = CASES[name]
    dim = c["dim"]
    whole, absorbing, reflecting = cases.mesh3d(3) if dim == 3 else cases.mesh2d(512)
    if c["kind"] == abi.REFLECTING_NONE:
        absorbing, reflecting = whole, None
    pts = cases.query_points(256, dim, 0.6, seed=17)
    n_walks = 64
    prob, out = str(tmp_path / "p.bin"), str(tmp_path / "o.bin")
    write_problem(prob, dim, c["estimator"], c["kind"], absorbing, reflecting, pts, n_walks, c["quantity"], c["family"],
                  c["params"], c["sigma"], c["mu"])
    r = subprocess.run([driver, prob, out], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stderr
    steps, rec, _ = read_result(out, len(pts))

    # the same problem straight through the C ABI
    if reflecting is None:
        sc = native.Scene(ctx, dim, absorbing=absorbing)
    else:
        sc = native.Scene(ctx, dim, **cases.scene_kwargs(dim, c["kind"], absorbing, reflecting, mu=c["mu"]))
    _, da, _, _ = sc.closest_point(abi.ABSORBING, pts)
    dr = sc.closest_point(abi.REFLECTING, pts)[1] if reflecting is not None else None
    hp = abi.HostSamplePoints(pts, da, dr, nWalks=n_walks, estimationQuantity=c["quantity"], seed=0xC0FFEE)
    hs = abi.HostStatistics(hp.n, dim)
    pde = abi.PdeDesc.manufactured(c["family"], c["params"], absorptionCoeff=c["sigma"], robinCoeff=c["mu"])
    steps_abi = sc.solve(c["estimator"], pde, abi.WalkSettings.default(), hp, hs)

    assert steps == steps_abi
    assert np.array_equal(rec[:, 0], hs.solutionMean)           # bit-identical: the mirror only marshals
    assert np.array_equal(rec[:, 2].astype(np.int32), hs.nSolutionEstimates)
    assert np.array_equal(rec[:, 4], hp.firstSphereRadius)
    assert np.array_equal(rec[:, 5:5 + dim], hs.gradientMean)
    # and the estimates are right: within 4 standard errors of the exact solution at >= 97 % of the points
    exact = cases.quadratic_exact(pts) if c["family"] == abi.SOL_QUADRATIC else np.exp(pts @ np.array(c["params"][:3], np.float32))
    se = np.sqrt(rec[:, 1]/np.maximum(1, rec[:, 2]))
    assert np.mean(np.abs(rec[:, 0] - exact) < 4.0*se + 2e-3) > 0.97
