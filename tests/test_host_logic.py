"""CPU tests of the host-side logic of the product: the wide quantised BVH builder (validated on the host,
no device), the procedural inputs and workload definitions, the byte accounting of the roofline."""
import numpy as np
import pytest

from tests import cases
from zombie_b200 import abi, native, procedural as P, workloads


@pytest.mark.parametrize("dim,reflecting", [(3, False), (3, True), (2, False), (2, True)])
def test_wide_bvh_is_sound(dim, reflecting):
    if dim == 3:
        pos, idx = P.bumpy_sphere(4)
    else:
        pos, idx = P.star_polygon(3000)
    r = native.host_bvh_check(dim, pos, idx, reflecting)
    assert r["violations"] == 0
    assert r["leaves"] >= len(idx)/4 and r["nodes"] >= r["leaves"]/8
    assert r["max_depth"] <= 24


def test_wide_bvh_ragged_and_degenerate_inputs():
    # a single primitive, coincident primitives, a degenerate (zero-area) triangle, a 5-primitive mesh
    pos = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]], np.float32)
    assert native.host_bvh_check(3, pos, np.array([[0, 1, 2]], np.int32))["violations"] == 0
    same = np.array([[0, 1, 2]]*9, np.int32)
    assert native.host_bvh_check(3, pos, same)["violations"] == 0
    degenerate = np.array([[0, 1, 1], [0, 1, 2], [1, 2, 3], [0, 2, 3], [0, 1, 3]], np.int32)
    assert native.host_bvh_check(3, pos, degenerate, True)["violations"] == 0
    pos2 = np.array([[0, 0], [1, 0], [1, 1]], np.float32)
    assert native.host_bvh_check(2, pos2, np.array([[0, 1], [1, 2], [2, 0]], np.int32), True)["violations"] == 0


def test_wide_bvh_rejects_bad_input():
    pos = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32)
    with pytest.raises(native.ZombieError, match="out of range"):
        native.host_bvh_check(3, pos, np.array([[0, 1, 3]], np.int32))
    bad = pos.copy()
    bad[1, 1] = np.nan
    with pytest.raises(native.ZombieError, match="non-finite"):
        native.host_bvh_check(3, bad, np.array([[0, 1, 2]], np.int32))
    with pytest.raises(native.ZombieError, match="empty"):
        native.host_bvh_check(3, pos, np.zeros((0, 3), np.int32))


def test_icosphere_is_watertight_and_outward():
    v, f = P.icosphere(3)
    assert len(f) == 20*4**3
    e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]])
    # every directed edge appears once and its reverse appears once
    key = set(map(tuple, e))
    assert len(key) == len(e) and all((b, a) in key for a, b in key)
    n = np.cross(v[f[:, 1]] - v[f[:, 0]], v[f[:, 2]] - v[f[:, 0]])
    assert ((n*v[f].mean(axis=1)).sum(axis=1) > 0).all()


def test_partition_mesh_compacts_vertices():
    pos, tri = P.bumpy_sphere(2)
    (ap, ai), (rp, ri) = P.partition_mesh(pos, tri, lambda c: c[:, 2] > 0)
    assert len(ai) + len(ri) == len(tri)
    assert ai.max() == len(ap) - 1 and ri.max() == len(rp) - 1
    assert (rp[ri].mean(axis=1)[:, 2] > 0).all() and (ap[ai].mean(axis=1)[:, 2] <= 0).all()


def test_workload_definitions_and_algorithmic_bytes():
    w = workloads.config3(subdivisions=3, n_points=100, n_walks=8)
    assert w.dim == 3 and w.estimator == abi.WALK_ON_STARS and w.quantity == abi.SOLUTION_AND_GRADIENT
    assert w.n_triangles(abi.ABSORBING) + w.n_triangles(abi.REFLECTING) == 1280
    # 1 M triangles -> depth 18 -> 1008 B closest point; Robin star radius 2016 B; ray 1008;
    # sphere sample 468; occlusion 1008
    class Big(workloads.Workload):
        def n_triangles(self, which):
            return 1 << 20
    big = Big.__new__(Big)
    big.__dict__.update(w.__dict__)
    assert big.algorithmic_bytes_per_step() == 1008 + 2016 + 1008 + 468 + 1008
    w1 = workloads.config1(256, 16, 4)
    assert w1.dim == 2 and w1.estimator == abi.WALK_ON_SPHERES and len(w1.points) > 50
    w2 = workloads.config2(256, 16, 4)
    assert w2.reflecting_kind == abi.REFLECTING_NEUMANN and w2.settings.splittingThreshold == pytest.approx(1.5)


def test_walk_settings_defaults_match_reference_short_constructor():
    """zombie::WalkSettings short constructor (point_estimation/common.h:21-40)"""
    s = abi.WalkSettings.default(1e-3, 2e-3, 77, True)
    assert s.silhouettePrecision == pytest.approx(1e-3) and s.russianRouletteThreshold == 0.0
    assert s.splittingThreshold == pytest.approx(abi.FLT_MAX) and s.maxWalkLength == 77
    assert s.stepsBeforeApplyingTikhonov == 0 and s.stepsBeforeUsingMaximalSpheres == 77
    assert s.solveDoubleSided == 1 and s.useGradientControlVariates == 1 and s.useGradientAntitheticVariates == 1
    assert s.useCosineSamplingForDerivatives == 0 and s.ignoreSourceContribution == 0 and s.printLogs == 0


def test_merge_statistics_equals_one_pass_welford():
    """sharding.merge_statistics == SampleStatistics::merge (common.h:268-310): merging the Welford states of two
    batches equals one Welford pass over their concatenation."""
    from zombie_b200 import abi, sharding
    rng = np.random.RandomState(5)
    n, dim = 17, 3

    def welford(samples):  # samples [n, k, ...]
        st_mean = np.zeros(samples.shape[:1] + samples.shape[2:], np.float64)
        st_m2 = np.zeros_like(st_mean)
        for j in range(samples.shape[1]):
            d = samples[:, j] - st_mean
            st_mean += d/(j + 1)
            st_m2 += d*(samples[:, j] - st_mean)
        return st_mean, st_m2

    def make(sol, grad):
        st = abi.HostStatistics(n, dim)
        m, m2 = welford(sol)
        st.solutionMean, st.solutionM2 = m.astype(np.float32), m2.astype(np.float32)
        g, g2 = welford(grad)
        st.gradientMean, st.gradientM2 = g.astype(np.float32), g2.astype(np.float32)
        st.nSolutionEstimates[:] = sol.shape[1]
        st.nGradientEstimates[:] = grad.shape[1]
        st.totalWalkLength[:] = 7*sol.shape[1]
        return st

    sa, sb = rng.normal(size=(n, 40)), rng.normal(size=(n, 24)) + 0.5
    ga, gb = rng.normal(size=(n, 40, dim)), rng.normal(size=(n, 24, dim))
    merged = sharding.merge_statistics(make(sa, ga), make(sb, gb))
    whole = make(np.concatenate([sa, sb], 1), np.concatenate([ga, gb], 1))
    assert np.allclose(merged.solutionMean, whole.solutionMean, atol=1e-5)
    assert np.allclose(merged.solutionM2, whole.solutionM2, rtol=1e-4)
    assert np.allclose(merged.gradientMean, whole.gradientMean, atol=1e-5)
    assert np.allclose(merged.gradientM2, whole.gradientM2, rtol=1e-4)
    assert np.array_equal(merged.nSolutionEstimates, whole.nSolutionEstimates)
    assert np.array_equal(merged.totalWalkLength, whole.totalWalkLength)
    # merging an empty batch is the identity
    empty = abi.HostStatistics(n, dim)
    same = sharding.merge_statistics(whole, empty)
    assert np.array_equal(same.solutionMean, whole.solutionMean) and np.array_equal(same.gradientM2, whole.gradientM2)
