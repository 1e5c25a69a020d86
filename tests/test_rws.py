"""Reverse Walk on Stars (BASELINE configs[3], second half): zb_rws_solve against the reference's
rws::ReverseWalkOnStarsSolver::solve (zombie/variance_reduction/reverse_walk_solver.h:287-330).

The estimator is Monte Carlo and the reference seeds from the clock, so parity is statistical: both must reach the exact
solution to the same accuracy on the same samples, and their six per-point running sums must agree within the spread
the reference shows between two of its own runs with different seeds.

CPU: the device code (walks in record-emitting mode + per-pair gather arithmetic) compiled for the host by
tools/host_rws.cu. GPU: the same through the C ABI."""
import os
import struct
import subprocess

import numpy as np
import pytest

from tests import cases
from tests.conftest import ROOT
from zombie_b200 import abi, bvc, procedural as P

OFFSET = 0.01


def problem(n_abs=12000, n_ref=6000, n_dom=6000, n_eval=160):
    """3D Poisson on the bumpy sphere: Dirichlet z <= 0, Neumann z > 0, u = x^2 - y^2/2 + z^2/4 + 0.3"""
    whole, absorbing, reflecting = cases.mesh3d(3)
    sa, sr = bvc.UniformBoundarySampler(*absorbing, seed=1), bvc.UniformBoundarySampler(*reflecting, seed=2)
    sd = bvc.UniformDomainSampler(lambda p: P.bumpy_sphere_inside(p), [-1.2]*3, [1.2]*3, P.bumpy_sphere_volume(), seed=3)
    sa.initialize(OFFSET, False)
    sr.initialize(0.0, False)
    s = bvc.Samples.concatenate([sa.generate_samples(n_abs, abi.ON_ABSORBING_BOUNDARY, OFFSET),
                                 sr.generate_samples(n_ref, abi.ON_REFLECTING_BOUNDARY, 0.0), sd.generate_samples(n_dom)])
    x = cases.query_points(n_eval, 3, 0.6, seed=5)
    return absorbing, reflecting, s, x


def reference_run(ref, absorbing, reflecting, s, x, seed):
    rs = ref.RefScene(3, **cases.scene_kwargs(3, abi.REFLECTING_NEUMANN, absorbing, reflecting))
    da = rs.closest_point(abi.ABSORBING, s.pt, want_normal=False)[1]
    dr = rs.closest_point(abi.REFLECTING, s.pt, want_normal=False)[1]
    smp = abi.HostRwsSamples(s.pt, s.normal, s.type, s.pdf, da, dr, seed=seed)
    ev = abi.HostRwsEvalPoints(x)
    counts, _ = rs.rws_solve(cases.quadratic_pde(), abi.WalkSettings.default(), smp, ev, OFFSET)
    assert np.array_equal(counts, smp.group_counts())
    return smp, ev, da, dr


def compare(totals, x, smp, ref_a, ref_b):
    """totals of the implementation under test against two reference runs"""
    ev = abi.HostRwsEvalPoints(x)
    ev.totals[:] = totals
    counts = smp.group_counts()
    exact = cases.quadratic_exact(x)
    err = ev.estimated_solution(counts) - exact
    err_a, err_b = ref_a.estimated_solution(counts) - exact, ref_b.estimated_solution(counts) - exact
    rmse, rmse_a, rmse_b = (float(np.sqrt(np.mean(e**2))) for e in (err, err_a, err_b))
    assert rmse < 1.5*max(rmse_a, rmse_b) + 0.01, (rmse, rmse_a, rmse_b)
    assert abs(err.mean()) < 0.05
    # each running sum, averaged over the evaluation points, within the run-to-run spread of the reference
    for col in (0, 1, 3, 5):
        ta, tb, t = ref_a.totals[:, col], ref_b.totals[:, col], ev.totals[:, col]
        spread = np.abs(ta - tb).mean() + 1e-3*np.abs(ta).mean()
        assert abs((t - 0.5*(ta + tb)).mean()) < 4.0*spread/np.sqrt(len(t)) + 0.02*np.abs(ta).mean(), col
        assert np.abs(t - 0.5*(ta + tb)).mean() < 3.0*spread, col
    assert not ev.totals[:, 2].any() and not ev.totals[:, 4].any()   # no normal-aligned samples in a single-sided solve


def test_device_reverse_walk_code_on_host_matches_reference(ref, tmp_path):
    absorbing, reflecting, s, x = problem()
    _, ref_a, da, dr = reference_run(ref, absorbing, reflecting, s, x, seed=0xF00D)
    smp, ref_b, _, _ = reference_run(ref, absorbing, reflecting, s, x, seed=0xD00F)
    exe = str(tmp_path / "host_rws")
    cmd = ["/usr/local/cuda/bin/nvcc", "-std=c++17", "-O2", "-DZB_HOST_STATS", "-Xcompiler", "-ffp-contract=off", "-o", exe,
           os.path.join(ROOT, "tools", "host_rws.cu"), os.path.join(ROOT, "zombie_b200", "csrc", "wide_bvh.cpp")]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout[-3000:]
    d = str(tmp_path)
    hdr = struct.pack("<4i16f4fQ", 3, abi.REFLECTING_NEUMANN, abi.SOL_QUADRATIC, 0, *([1.0, -0.5, 0.25, 0.3] + [0.0]*12), 0.0, 0.0,
                      OFFSET, 0.0, 0xBEEF)
    open(os.path.join(d, "hdr.bin"), "wb").write(hdr)
    for name, arr, dt in (("apos", absorbing[0], np.float32), ("aidx", absorbing[1], np.int32), ("rpos", reflecting[0], np.float32),
                          ("ridx", reflecting[1], np.int32), ("spt", s.pt, np.float32), ("snormal", s.normal, np.float32),
                          ("stype", s.type, np.int32), ("spdf", s.pdf, np.float32), ("ept", x, np.float32)):
        np.ascontiguousarray(arr, dt).tofile(os.path.join(d, name + ".bin"))
    r = subprocess.run([exe, d], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0 and "status 0" in r.stdout, r.stdout
    totals = np.fromfile(os.path.join(d, "totals.bin"), np.float32).reshape(-1, 6)
    compare(totals, x, smp, ref_a, ref_b)


@pytest.mark.gpu
def test_gpu_reverse_walk_matches_reference(ctx, ref):
    from zombie_b200 import native
    absorbing, reflecting, s, x = problem()
    _, ref_a, da, dr = reference_run(ref, absorbing, reflecting, s, x, seed=0xF00D)
    smp_ref, ref_b, _, _ = reference_run(ref, absorbing, reflecting, s, x, seed=0xD00F)
    sc = native.Scene(ctx, 3, **cases.scene_kwargs(3, abi.REFLECTING_NEUMANN, absorbing, reflecting))
    gda = sc.closest_point(abi.ABSORBING, s.pt, want_normal=False)[1]
    gdr = sc.closest_point(abi.REFLECTING, s.pt, want_normal=False)[1]
    assert np.allclose(gda, da, rtol=1e-6) and np.allclose(gdr, dr, rtol=1e-6, atol=1e-9)
    smp = abi.HostRwsSamples(s.pt, s.normal, s.type, s.pdf, gda, gdr, seed=0xBEEF)
    state_before = smp.rngState.copy()
    ev = abi.HostRwsEvalPoints(x)
    steps, records = sc.rws_solve(cases.quadratic_pde(), abi.WalkSettings.default(), smp, ev, OFFSET)
    assert steps > 0 and records >= steps          # one record per step, plus the steps that ended a walk
    assert (smp.rngState != state_before).all()    # generators advance, so a second solve draws fresh walks
    compare(ev.totals, x, smp, ref_a, ref_b)
    # progressive: a second solve accumulates onto the same totals (reverse_walk_solver.h: totals are running sums)
    first = ev.totals.copy()
    sc.rws_solve(cases.quadratic_pde(), abi.WalkSettings.default(), smp, ev, OFFSET)
    assert (ev.totals[:, 1] > first[:, 1]).mean() > 0.95 and abs(ev.totals[:, 1].mean()/first[:, 1].mean() - 2.0) < 0.5
    # evaluation points on the absorbing boundary take no splats
    ev2 = abi.HostRwsEvalPoints(x[:8], type=abi.ON_ABSORBING_BOUNDARY)
    sc.rws_solve(cases.quadratic_pde(), abi.WalkSettings.default(), smp, ev2, OFFSET)
    assert not ev2.totals.any()


@pytest.mark.gpu
def test_gpu_bvc_splat_2d_yukawa(ctx, ref):
    """the 2D screened-Poisson free-space kernels (modified Bessel functions, distributions.h:215-262) of zb_bvc_splat;
    here, after the longer-standing GPU tests, because it first runs on a GPU at the end of round 1"""
    from tests import test_bvc
    test_bvc.check_gpu_splat(ctx, ref, dict(dim=2, sigma=3.0, reg=0.0, robin_cutoff=1.0))
