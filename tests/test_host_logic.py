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
