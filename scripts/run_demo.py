"""Runs one of the reference's 2D demo configs (demo/model_problems/*/*.json: solverType wos or wost, channels 1 or 4, single- or
double-sided, PFM or PNG textures) through zombie_b200's Python surface with the call sequence of demo/demo.cpp /
demo/model_problem.h, and writes the solution image (.npz with the float image, its per-pixel variance and the mask; plus the
.pfm / .png the demo would write).

    python scripts/run_demo.py <config.json> <root> <out prefix>

<root> is the directory the config's paths are relative to (the reference checkout, or a copy of the files it names). The
tests run the engine configs from a committed fixture (tests/test_demo_engine.py); this script is for the configs whose inputs
are too large to commit (nautilus: 54,590 segments, 2048^2 colour textures, 1024^2 grid, four channels, two-sided Dirichlet
data) and for timing them."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zombie_b200 import abi, pfm, zombie  # noqa: E402


def load_texture(path, channels):
    """Image<C>::read (demo/image.h:163-262) in buffer order; C = 1 for T = float, 3 for T = Array4 (image.h:31-41)"""
    if path.endswith(".pfm"):
        rgb = pfm.read_pfm(path)
        if rgb.shape[2] == 1:
            rgb = np.repeat(rgb, 3, axis=2)
    else:
        from PIL import Image
        # readPNG sets stbi_set_flip_vertically_on_load (image.h:243): buffer row 0 is the BOTTOM row of the picture
        rgb = (np.asarray(Image.open(path).convert("RGB"), np.float32)[::-1]/np.float32(255.0)).astype(np.float32)
    if channels == 1:
        return pfm.luminance(rgb)
    return np.concatenate([rgb, np.zeros_like(rgb[..., :1])], axis=-1)      # ModelProblem::getPDEValue: head<3>() of a zero Array4


def dense(buffer):
    """texture in buffer order -> DenseGrid data ([nx, ny(, channels)], y fastest): Image::get(uv) reads row h - 1 - int(uv.y h)"""
    return np.ascontiguousarray(np.swapaxes(buffer[::-1], 0, 1))


def main():
    config_path, root, out = sys.argv[1], sys.argv[2], sys.argv[3]
    cfg = json.load(open(config_path))
    sc, mp, oc = cfg["solver"], cfg["modelProblem"], cfg["output"]
    channels = mp.get("channels", 1)
    two_sided = mp.get("solveDoubleSided", False)
    dim = 2
    t0 = time.time()
    path = lambda key: os.path.join(root, mp[key])

    positions, indices = zombie.FloatNList(dim=dim), zombie.IntNList(dim=dim)
    zombie.Utils.load_boundary_mesh(path("geometry"), positions, indices, dim=dim)
    if mp.get("normalizeDomain", True):
        zombie.Utils.normalize(positions, dim=dim)
    if mp.get("flipOrientation", True):
        zombie.Utils.flip_orientation(indices, dim=dim)
    bbox = zombie.Utils.compute_bounding_box(positions, True, 1.0, dim=dim)

    def grid_callback(factory, key, aligned_key=None):
        tex = load_texture(path(key), channels)
        shape = [tex.shape[1], tex.shape[0]]
        data = dense(tex).reshape(-1, channels) if channels > 1 else dense(tex).ravel()
        if aligned_key and two_sided:
            al = load_texture(path(aligned_key), channels)
            al = dense(al).reshape(-1, channels) if channels > 1 else dense(al).ravel()
            return factory(data, al, shape, bbox[0], bbox[1], dim=dim, channels=channels)
        return factory(data, shape, bbox[0], bbox[1], dim=dim, channels=channels)

    pde = zombie.Core.PDE(dim=dim, channels=channels)
    pde.source = grid_callback(zombie.Utils.get_dense_grid_source_callback, "sourceValue")
    pde.dirichlet = grid_callback(zombie.Utils.get_dense_grid_dirichlet_callback, "absorbingBoundaryValue", "absorbingBoundaryNormalAlignedValue")
    pde.robin = grid_callback(zombie.Utils.get_dense_grid_robin_callback, "reflectingBoundaryValue", "reflectingBoundaryNormalAlignedValue")
    mu = mp.get("robinCoeff", 0.0)
    pde.robin_coeff = zombie.Core.get_constant_robin_coefficient_callback(mu, dim=dim)
    refl = pfm.luminance(load_texture(path("isReflectingBoundary"), 4)[..., :3])
    pde.has_reflecting_boundary_conditions = zombie.Utils.get_dense_grid_indicator_callback(dense(refl).ravel(), [refl.shape[1], refl.shape[0]],
                                                                                           bbox[0], bbox[1], dim=dim)
    pde.absorption_coeff = mp.get("absorptionCoeff", 0.0)
    pde.is_source_constant = False
    pde.are_robin_conditions_pure_neumann = mu == 0.0
    pde.are_robin_coeffs_nonnegative = mu >= 0.0

    ap, ai, rp, ri = (zombie.FloatNList(dim=dim), zombie.IntNList(dim=dim), zombie.FloatNList(dim=dim), zombie.IntNList(dim=dim))
    zombie.Utils.partition_boundary_mesh(pde.has_reflecting_boundary_conditions, positions, indices, ap, ai, rp, ri, dim=dim)
    queries = zombie.Core.GeometricQueries(mp.get("domainIsWatertight", False), bbox[0], bbox[1], dim=dim)
    if len(ai):
        h = zombie.Utils.FcpwDirichletBoundaryHandler(dim=dim)
        h.build_acceleration_structure(ap, ai)
        zombie.Utils.populate_geometric_queries_for_dirichlet_boundary(h, queries, dim=dim)
    if len(ri):
        h = zombie.Utils.FcpwNeumannBoundaryHandler(dim=dim)
        h.build_acceleration_structure(rp, ri, zombie.Utils.get_ignore_candidate_silhouette_callback(two_sided))
        zombie.Utils.populate_geometric_queries_for_neumann_boundary(h, zombie.Utils.get_branch_traversal_weight_callback(), queries, dim=dim)
    scene = queries.scene()
    t_setup = time.time() - t0

    # createGridPoints, computeDistanceInfo, createSamplePoints (grid.h:22-40, demo.cpp:9-46), batched
    res = oc["gridRes"]
    ii, jj = np.meshgrid(np.arange(res), np.arange(res), indexing="ij")
    extent = bbox[1] - bbox[0]
    grid = np.stack([(ii.ravel()/np.float32(res))*extent[0] + bbox[0][0], (jj.ravel()/np.float32(res))*extent[1] + bbox[0][1]], axis=1).astype(np.float32)
    da = scene.closest_point(abi.ABSORBING, grid, want_normal=False)[1] if len(ai) else np.full(len(grid), abi.FLT_MAX, np.float32)
    dr = scene.closest_point(abi.REFLECTING, grid, want_normal=False)[1] if len(ri) else np.full(len(grid), abi.FLT_MAX, np.float32)
    valid = scene.inside_domain(grid).astype(bool) | bool(two_sided)
    where = np.where(valid)[0]

    eps_a = sc.get("epsilonShellForAbsorbingBoundary", 1e-3)
    ignore = [sc.get(k, False) for k in ("ignoreAbsorbingBoundaryContribution", "ignoreReflectingBoundaryContribution", "ignoreSourceContribution")]
    rr, split, length = sc.get("russianRouletteThreshold", 0.0), sc.get("splittingThreshold", np.inf), sc.get("maxWalkLength", 1024)
    if cfg["solverType"] == "wos":      # runWalkOnSpheres (demo.cpp:49-92)
        estimator = abi.WALK_ON_SPHERES
        settings = zombie.Solvers.WalkSettings(eps_a, 0.0, 0.0, rr, split, length, 0, 0, two_sided, True, True, False, ignore[0], True, ignore[2], False)
    elif cfg["solverType"] == "wost":   # runWalkOnStars (demo.cpp:95-150)
        estimator = abi.WALK_ON_STARS
        settings = zombie.Solvers.WalkSettings(eps_a, sc.get("epsilonShellForReflectingBoundary", 1e-3), 1e-3, rr, split, length, 0, length, two_sided,
                                               True, True, False, ignore[0], ignore[1], ignore[2], False)
    else:
        raise SystemExit("solverType %s: see tests/test_demo_engine.py for the bvc / rws flows" % cfg["solverType"])
    n_walks = sc.get("nWalks", 128)
    pts = abi.HostSamplePoints(grid[where], da[where], dr[where], nWalks=n_walks, estimationQuantity=abi.SOLUTION, seed=0xD370)
    stats = [abi.HostStatistics(len(where), dim) for _ in range(channels)]
    t1 = time.time()
    descs = pde.descriptors()
    steps = scene.solve_channels(estimator, descs, settings, pts, stats) if channels > 1 else scene.solve(estimator, descs[0], settings, pts, stats[0])
    t_solve = time.time() - t1

    # saveGridValues (grid.h:43-85)
    keep = np.minimum(da, dr)[where] >= oc.get("boundaryDistanceMask", 0.0)
    i, j = np.divmod(where[keep], res)
    image = np.zeros((res, res, channels), np.float32)
    variance = np.zeros((res, res, channels), np.float32)
    visible = np.zeros((res, res), bool)
    for c in range(channels):
        image[j, i, c] = stats[c].solutionMean[keep]
        variance[j, i, c] = stats[c].solution_variance()[keep]
    visible[j, i] = True
    np.savez_compressed(out + ".npz", image=image, variance=variance, visible=visible, n_walks=n_walks)
    shown = image[..., :3] if channels > 1 else image[..., 0]
    pfm.write_pfm(out + ".pfm", shown)
    from PIL import Image
    png = np.clip((np.clip(shown, 0.0, 1.0)*255.0).astype(np.int32), 0, 255).astype(np.uint8)     # Image::writePNG (image.h:296-310)
    Image.fromarray(png).save(out + ".png")
    print(json.dumps({"config": os.path.basename(config_path), "segments": len(indices), "grid": res, "channels": channels, "points": int(len(where)),
                      "walks": int(len(where))*n_walks, "walk_steps": int(steps), "setup_s": round(t_setup, 2), "solve_s": round(t_solve, 2),
                      "steps_per_s": steps/t_solve}))


if __name__ == "__main__":
    main()
