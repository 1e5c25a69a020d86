"""Build time of the acceleration structures, host SAH builder vs on-device builder (zombie_b200/csrc/bvh_build.cuh), and the
rate of the batched queries over either hierarchy. Usage: python scripts/build_bench.py [subdivisions=8] [queries=200000]
ZB_BUILD_TRACE=1 prints the device builder's stages."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from zombie_b200 import abi, native, workloads  # noqa: E402

sub = int(sys.argv[1]) if len(sys.argv) > 1 else 8
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 200000
w = workloads.config3(subdivisions=sub, n_points=nq, n_walks=1)
kw = w.scene_kwargs()
ctx = native.Context(0)
out = {"triangles": len(kw["absorbing"][1]) + len(kw["reflecting"][1])}
scenes = {}
for name, builder in (("host_sah", abi.BVH_BUILDER_HOST_SAH), ("device", abi.BVH_BUILDER_DEVICE), ("device_again", abi.BVH_BUILDER_DEVICE)):
    t = time.perf_counter()
    s = native.Scene(ctx, 3, builder=builder, **kw)
    wall = time.perf_counter() - t
    prof = [s.build_profile(b) for b in (abi.ABSORBING, abi.REFLECTING)]
    out[name] = {"create_wall_ms": 1e3*wall, "absorbing_ms": prof[0]["ms"], "reflecting_ms": prof[1]["ms"],
                 "nodes": [s.stats(b)["nodes"] for b in (0, 1)], "depth": [s.stats(b)["max_depth"] for b in (0, 1)]}
    scenes[name] = s
# moving geometry: every vertex moves, the hierarchy is rebuilt on the device
s = scenes["device"]
for b, key in ((abi.ABSORBING, "absorbing"), (abi.REFLECTING, "reflecting")):
    pos = (kw[key][0]*1.01).astype(np.float32)
    t = time.perf_counter()
    s.update_vertices(b, pos)
    out.setdefault("update_vertices", {})[key] = {"wall_ms": 1e3*(time.perf_counter() - t), "build_ms": s.build_profile(b)["ms"]}
x = w.points
minr = np.full(len(x), 1e-3, np.float32)
for name in ("host_sah", "device_again"):
    s = scenes[name]
    _, da, _, _ = s.closest_point(abi.ABSORBING, x)
    r = {}
    for q, fn in (("closest_point", lambda: s.closest_point(abi.REFLECTING, x)), ("star_radius", lambda: s.star_radius(x, minr, da))):
        fn()
        t = time.perf_counter()
        fn()
        r[q + "_Mq_per_s"] = 1e-6*len(x)/(time.perf_counter() - t)
    out[name]["queries"] = r
print(json.dumps(out))
