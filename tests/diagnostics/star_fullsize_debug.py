"""Diagnostic: star-radius rows of the full-size parity test that deviate from the reference's BVH; prints every value."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from tests import cases
from zombie_b200 import abi, native, workloads
from oracle import ref

nq = 20000
w = workloads.config3(subdivisions=8, n_points=256, n_walks=16)
kw = w.scene_kwargs()
ctx = native.Context(0)
rs, gs = ref.RefScene(3, **kw), native.Scene(ctx, 3, **kw)
x = cases.query_points(nq, 3, 0.95, seed=141)
_, da, _, _ = rs.closest_point(abi.ABSORBING, x)
minr = np.full(nq, 1e-3, np.float32)
rr = rs.star_radius(x, minr, da)
runs = [gs.star_radius(x, minr, da) for _ in range(2)]
gr = runs[0]
print("repeatable:", np.array_equal(runs[0], runs[1]))
# one query per call: no other query shares the warp
bad = np.abs(rr - gr) > 1e-3*max(1.0, float(rr.max()))
rb = ref.RefScene(3, use_baseline=True, **kw)
base = rb.star_radius(x[bad], minr[bad], da[bad])
for k, i in enumerate(np.where(bad)[0]):
    single = gs.star_radius(x[i:i + 1], minr[i:i + 1], da[i:i + 1])[0]
    print(i, x[i].tolist(), "da %.9g ref %.9g gpu %.9g gpu-alone %.9g baseline %.9g" % (da[i], rr[i], gr[i], single, base[k]))
np.savez(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/star_full.npz", x=x, da=da, rr=rr, gr=gr, bad=np.where(bad)[0], base=base)
