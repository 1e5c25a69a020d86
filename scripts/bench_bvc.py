#!/usr/bin/env python
"""Throughput of the Boundary Value Caching splat (BASELINE configs[3]): N_eval evaluation
points x N_samples cache samples through zb_bvc_splat_dev, evaluation points resident in HBM. Reports pairs/s,
FP32 TFLOP/s at SURVEY §8d's 80 flop per pair against the non-tensor FP32 peak (#SM x 128 x 2 x SM clock), and the
reference's TBB splat on a bounded sample of the same evaluation points.

    python scripts/bench_bvc.py [--eval 10000000] [--samples 12288] [--sigma 0] [--steps 3]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from zombie_b200 import abi, native  # noqa: E402


def make(n_samples, n_eval, seed=3):
    rng = np.random.RandomState(seed)
    y = rng.normal(size=(n_samples, 3))
    y /= np.linalg.norm(y, axis=1, keepdims=True)
    typ = np.repeat([abi.ON_ABSORBING_BOUNDARY, abi.ON_REFLECTING_BOUNDARY, abi.IN_DOMAIN, abi.IN_DOMAIN],
                    [n_samples//4, n_samples//4, n_samples//4, n_samples - 3*(n_samples//4)])
    y[typ == abi.IN_DOMAIN] *= rng.uniform(0, 1, (int((typ == abi.IN_DOMAIN).sum()), 1))**(1/3)
    s = abi.HostBvcSamples(y, y/np.maximum(1e-6, np.linalg.norm(y, axis=1, keepdims=True)), typ, np.full(n_samples, 0.08),
                           rng.normal(size=n_samples), rng.normal(size=n_samples), rng.normal(size=n_samples),
                           np.zeros(n_samples))
    x = rng.uniform(-0.55, 0.55, (n_eval, 3)).astype(np.float32)
    dist = (1.0 - np.linalg.norm(x, axis=1)).astype(np.float32)
    return s, x, dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--eval", type=int, default=10_000_000)
    ap.add_argument("--samples", type=int, default=12288)
    ap.add_argument("--sigma", type=float, default=0.0)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--ref-eval", type=int, default=20000)
    args = ap.parse_args()
    import torch
    dev = torch.device("cuda", 0)
    ctx = native.Context(0)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    s, x, dist = make(args.samples, args.eval)
    n = args.eval
    t = {"pt": torch.from_numpy(x).to(dev), "type": torch.zeros(n, dtype=torch.int32, device=dev),
         "da": torch.from_numpy(dist).to(dev), "dr": torch.from_numpy(dist).to(dev),
         "cnt": torch.zeros((n, 5), dtype=torch.int32, device=dev), "sum": torch.zeros((n, 5, 4), dtype=torch.float32, device=dev),
         "sol": torch.zeros(n, dtype=torch.float32, device=dev), "grad": torch.zeros((n, 3), dtype=torch.float32, device=dev)}
    es = abi.BvcEvalPoints(n, *[t[k].data_ptr() for k in ("pt", "type", "da", "dr", "cnt", "sum", "sol", "grad")])
    ms = []
    for it in range(args.steps + 1):
        t["cnt"].zero_(); t["sum"].zero_()
        pairs = ctx.bvc_splat(3, args.sigma, s, None, device_eval_struct=es)
        torch.cuda.synchronize()
        if it > 0:
            ms.append(ctx.last_splat_ms())
    best = float(np.median(ms))
    sm = ctx.sm_count
    clock = 1.965e9
    try:
        import subprocess
        clock = float(subprocess.run(["nvidia-smi", "--query-gpu=clocks.max.sm", "--format=csv,noheader,nounits"],
                                     stdout=subprocess.PIPE, text=True).stdout.split()[0])*1e6
    except Exception:
        pass
    peak = sm*128*2*clock/1e12
    flops = pairs*80.0
    line = {"metric": "bvc splat pairs/sec", "value": pairs/(best*1e-3), "unit": "pairs/s", "ms": best, "n_eval": n, "n_samples": args.samples,
            "sigma": args.sigma, "roofline": {"bound": "fp32", "achieved": flops/(best*1e-3)/1e12, "peak": peak, "unit": "TFLOP/s",
                                              "frac": flops/(best*1e-3)/1e12/peak, "flop_per_pair": 80, "sm": sm}}
    # cpu_baseline leg (like bench.py's): the reference's splat on a bounded sample of the same evaluation points, timed on the host;
    # its values double as a spot check of the device result
    from oracle import ref
    if ref.available():
        k = min(args.ref_eval, n)
        e = abi.HostBvcEvalPoints(x[:k], dist[:k], dist[:k])
        sol, grad, cnt, mean, secs = ref.bvc_splat(3, args.sigma, s, e)
        gsol = t["sol"][:k].cpu().numpy()
        line["cpu_baseline"] = {"value": k*args.samples/secs, "unit": "pairs/s", "cores": ref.hardware_threads(), "kind": "reference",
                                "sample": "%d evaluation points x %d samples (%.1f s)" % (k, args.samples, secs)}
        line["max_abs_diff_vs_reference"] = float(np.abs(gsol - sol).max())
        line["solution_scale"] = float(np.abs(sol).max())
    print(json.dumps(line))


if __name__ == "__main__":
    main()
