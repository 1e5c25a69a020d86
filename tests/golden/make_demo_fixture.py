"""Generates tests/golden/demo_engine.npz from the reference's 2D demo model problem (run in a container that has /root/reference):
demo/model_problems/engine/data/{geometry.obj, *_value.pfm, is_reflecting_boundary.pfm} — the inputs of engine/{wost,bvc,rws}.json —
and the solutions the reference committed for them, demo/model_problems/engine/solutions/{wost,bvc,rws}.pfm (the only end-to-end
artefacts of the estimators the reference holds). Images are stored in the buffer order of the demo's Image class
(zombie_b200/pfm.py); the solutions keep one of their three identical channels.
    python tests/golden/make_demo_fixture.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from zombie_b200 import pfm  # noqa: E402

SRC = "/root/reference/demo/model_problems/engine/"
out = {}
for key, name in (("absorbing", "data/absorbing_boundary_value"), ("reflecting", "data/reflecting_boundary_value"),
                  ("source", "data/source_value"), ("is_reflecting", "data/is_reflecting_boundary")):
    out[key] = pfm.read_pfm(SRC + name + ".pfm")
for key in ("wost", "bvc", "rws"):
    img = pfm.read_pfm(SRC + "solutions/%s.pfm" % key)
    assert np.array_equal(img[..., 0], img[..., 1]) and np.array_equal(img[..., 1], img[..., 2])
    out["ref_" + key] = img[..., 0]
out["obj"] = np.frombuffer(open(SRC + "data/geometry.obj", "rb").read(), dtype=np.uint8)
for name in ("wost", "bvc", "rws"):
    out["config_" + name] = np.frombuffer(open(SRC + name + ".json", "rb").read(), dtype=np.uint8)
dst = os.path.join(ROOT, "tests", "golden", "demo_engine.npz")
np.savez_compressed(dst, **out)
print(dst, os.path.getsize(dst), "bytes")
