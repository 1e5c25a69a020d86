#!/usr/bin/env python
"""Benchmark of the walk hot path: BASELINE.json's metric "walk steps/sec" on the headline workload C3
(3D screened Poisson, Robin + Dirichlet boundary of a 1,310,720-triangle bumpy icosphere, WalkOnStars
with gradient estimates, 1,048,576 evaluation points x 64 walks).

A *step* is one zb_solve over one batch of --batch evaluation points of that workload (a different,
contiguous slice of the 1M points each step). `value` is whole-job walk steps/s with the points already
resident in HBM (zb_solve_dev); `e2e` is the same through the host-buffer C-ABI call zb_solve with the
host<->device copies of every in/out array inside the timed region.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]

N > 1: launched under torchrun, one rank per GPU; the evaluation points are sharded over the ranks (the
BVH is replicated), there is no data-path collective, timing is the max over ranks.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from zombie_b200 import abi, workloads  # noqa: E402


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f).get("hbm_gbs", 6650.0), "measured"
    return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.stop_flag = threading.Event()
        self.max_mhz = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=5).stdout.strip()
                f = [s.strip() for s in out.split(",")]
                self.samples.append(float(f[0]))
                self.max_mhz = float(f[1])
                for name, v in zip(names, f[2:]):
                    if v.lower().startswith("active"):
                        self.reasons.add(name)
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def make_points(w, idx, dist_a, dist_r, seed_base):
    return abi.HostSamplePoints(w.points[idx], dist_a[idx], dist_r[idx], nWalks=w.n_walks, estimationQuantity=w.quantity,
                                streams=np.asarray(idx, np.uint64) + np.uint64(seed_base))


# ------------------------------------------------------------------------------------------------
# reference arm: the unmodified reference (oracle/_ref, built from /root/reference by oracle/Makefile)
# ------------------------------------------------------------------------------------------------
def ref_solve_sample(scene, w, idx):
    """one WalkOnStars::solve of the reference (TBB default arena) on the evaluation points idx; (steps, seconds)"""
    _, da, _, _ = scene.closest_point(abi.ABSORBING, w.points[idx])
    _, dr, _, _ = scene.closest_point(abi.REFLECTING, w.points[idx])
    pts = abi.HostSamplePoints(w.points[idx], da, dr, nWalks=w.n_walks, estimationQuantity=w.quantity, streams=idx.astype(np.uint64))
    st = abi.HostStatistics(len(idx), 3)
    return scene.solve(w.estimator, w.pde, w.settings, pts, st, threads=0)


def ref_sample_size(scene, w, target_s, cap):
    """number of evaluation points the reference solves in about target_s seconds on this host (pilot run)"""
    rng = np.random.RandomState(321)
    n0 = 64
    _, secs = ref_solve_sample(scene, w, np.sort(rng.choice(len(w.points), n0, replace=False)))
    return int(min(cap, max(n0, n0*target_s/max(secs, 1e-3))))


def run_reference(args, rank, world):
    if rank != 0:
        return
    from oracle import ref
    if not ref.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libzombie_ref.so has not been built"}))
        return
    w = workloads.config3(args.subdiv, args.points, args.walks)
    scene = ref.RefScene(3, **w.scene_kwargs())
    cores = ref.hardware_threads()
    n_runs = args.warmup + args.steps
    n = args.ref_points or ref_sample_size(scene, w, min(20.0, 150.0/max(1, n_runs)), 16384)
    rng = np.random.RandomState(123)
    total_steps, total_secs = 0, 0.0
    for s in range(n_runs):
        idx = np.sort(rng.choice(len(w.points), n, replace=False))
        steps, secs = ref_solve_sample(scene, w, idx)
        if s >= args.warmup:
            total_steps += steps
            total_secs += secs
    value = total_steps/total_secs
    sample = ("%d random evaluation points x %d walks of the workload per step, TBB default arena (%d threads), -O3 -DNDEBUG"
              % (n, w.n_walks, cores))
    print(json.dumps({
        "impl": "reference", "metric": "walk steps/sec", "value": value, "unit": "steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3*total_secs/args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": w.name, "sample": sample},
        "cpu_baseline": {"value": value, "unit": "steps/s", "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


# ------------------------------------------------------------------------------------------------
# native arm
# ------------------------------------------------------------------------------------------------
def run_native(args, rank, world, local_rank):
    import torch
    from zombie_b200 import native

    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    w = workloads.config3(args.subdiv, args.points, args.walks)
    ctx = native.Context(local_rank)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    t0 = time.time()
    scene = native.Scene(ctx, 3, **w.scene_kwargs())
    build_s = time.time() - t0

    # this rank's shard of the evaluation points (interleaved for load balance); distances to the two
    # boundaries are part of SamplePoint construction in the reference (demo.cpp:9-46) -> setup, untimed
    shard = np.arange(rank, len(w.points), world)
    n_steps = args.warmup + args.steps
    _, dist_a, _, _ = scene.closest_point(abi.ABSORBING, w.points, want_normal=False)
    _, dist_r, _, _ = scene.closest_point(abi.REFLECTING, w.points, want_normal=False)

    def to_dev(a):
        return torch.from_numpy(np.ascontiguousarray(a)).to(dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    STAT_KEYS = ("solutionMean", "solutionM2", "gradientMean", "gradientM2", "totalFirstSourceContribution",
                 "totalDerivativeContribution", "nSolutionEstimates", "nGradientEstimates", "totalWalkLength", "totalSplits")
    POINT_KEYS = ("pt", "normal", "directionForDerivative", "type", "estimationQuantity", "distToAbsorbingBoundary",
                  "distToReflectingBoundary", "estimateBoundaryNormalAligned", "nWalks", "rngState", "rngInc", "firstSphereRadius")

    def device_set(idx, seed_base):
        hp = make_points(w, idx, dist_a, dist_r, seed_base)
        hs = abi.HostStatistics(hp.n, 3)
        tensors = {}
        for k in POINT_KEYS:
            a = getattr(hp, k)
            tensors[k] = to_dev(a.view(np.int64) if a.dtype == np.uint64 else a)
        for k in STAT_KEYS:
            tensors[k] = to_dev(getattr(hs, k))
        tensors["steps"] = torch.zeros(1, dtype=torch.int64, device=dev)
        ps = abi.SamplePoints(hp.n, *[tensors[k].data_ptr() for k in POINT_KEYS])
        ss = abi.Statistics(*[tensors[k].data_ptr() for k in STAT_KEYS])
        return ps, ss, tensors

    # ---- batch size: the points one step solves. --batch 0 sizes it with a pilot solve so that a step takes
    # about --step-seconds on this device (the persistent walk kernel wants >> 1 walk per lane, but the default
    # run has to end within minutes whatever the kernel's speed)
    batch = args.batch
    pilot = None
    if batch <= 0:
        n_pilot = min(2048, len(shard))
        ps, ss, tensors = device_set(shard[-n_pilot:], 7 << 40)
        t0 = time.perf_counter()
        scene.solve_dev(w.estimator, w.pde, w.settings, ps, ss, tensors["steps"].data_ptr())
        torch.cuda.synchronize()
        pilot_s = time.perf_counter() - t0
        per_point_s = pilot_s/n_pilot
        batch = int(args.step_seconds/max(per_point_s, 1e-9))
        batch = 1 << max(13, min(16, int(math.floor(math.log2(max(batch, 1))))))
        pilot = {"points": n_pilot, "seconds": pilot_s}
    if dist is not None:  # every rank must use the same batch
        t = torch.tensor([batch], dtype=torch.int64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        batch = int(t.item())
    batch = min(batch, len(shard))

    def batch_indices(s):
        start = (s*batch) % max(1, len(shard) - batch + 1)
        return shard[start:start + batch]

    # ---- device-resident arm --------------------------------------------------------------------
    dev_steps = [device_set(batch_indices(s), 0) for s in range(n_steps)]

    def run_dev(s):
        ps, ss, tensors = dev_steps[s]
        scene.solve_dev(w.estimator, w.pde, w.settings, ps, ss, tensors["steps"].data_ptr())

    for s in range(args.warmup):
        run_dev(s)
    barrier()
    launches_before = ctx.last_solve_profile()["launches"] if args.warmup > 0 else 0
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    walk_ms, walk_tasks, grid = 0.0, 0, 0
    ev0.record(stream)
    for s in range(args.warmup, n_steps):
        run_dev(s)
        # the library brackets the walk kernel of every solve with CUDA events on the launching stream; reading
        # them costs nothing here because zb_solve_dev has already waited for the solve's task count
        prof = ctx.last_solve_profile()
        walk_ms += prof["walk_ms"]
        walk_tasks += prof["walks"]
        grid = prof["walk_grid"]
    ev1.record(stream)
    barrier()
    dev_ms = ev0.elapsed_time(ev1)
    clocks = sampler.summary()
    launches_timed = ctx.last_solve_profile()["launches"] - launches_before
    steps_dev = int(sum(int(dev_steps[s][2]["steps"].item()) for s in range(args.warmup, n_steps)))
    prof_steps, local_walk_ms = steps_dev, walk_ms  # this rank's kernel: its steps / its kernel time

    # ---- end-to-end arm: host buffers through zb_solve --------------------------------------------
    # (at most 3 timed steps after 1 warm-up: the end-to-end arm repeats the device arm's work plus the copies, and the
    # whole default run has to end within minutes)
    e2e_warmup, e2e_steps = min(1, args.warmup), min(3, args.steps)
    host_sets = []
    for s in range(e2e_warmup + e2e_steps):
        hp = make_points(w, batch_indices(s), dist_a, dist_r, 1 << 40)
        hs = abi.HostStatistics(hp.n, 3)
        # pinned host memory
        for obj, keys in ((hp, ("pt", "normal", "directionForDerivative", "type", "estimationQuantity", "distToAbsorbingBoundary",
                                "distToReflectingBoundary", "estimateBoundaryNormalAligned", "nWalks", "rngState", "rngInc",
                                "firstSphereRadius")),
                          (hs, ("solutionMean", "solutionM2", "gradientMean", "gradientM2", "totalFirstSourceContribution",
                                "totalDerivativeContribution", "nSolutionEstimates", "nGradientEstimates", "totalWalkLength",
                                "totalSplits"))):
            for k in keys:
                a = getattr(obj, k)
                v = a.view(np.int64) if a.dtype == np.uint64 else a
                t = torch.from_numpy(v.copy()).pin_memory()
                pinned = t.numpy()
                setattr(obj, k, pinned.view(np.uint64) if a.dtype == np.uint64 else pinned)
                obj.__dict__.setdefault("_keep", []).append(t)
        host_sets.append((hp, hs))
    n, dim = host_sets[0][0].n, 3
    h2d = n*(4*dim*3 + 4*6 + 1 + 8*2) + n*(4*2 + 4*dim*2 + 4*2 + 4*4)
    d2h = n*(8 + 4) + n*(4*2 + 4*dim*2 + 4*2 + 4*4) + 8
    for s in range(e2e_warmup):
        scene.solve(w.estimator, w.pde, w.settings, *host_sets[s])
    barrier()
    t0 = time.perf_counter()
    steps_e2e = 0
    for s in range(e2e_warmup, e2e_warmup + e2e_steps):
        steps_e2e += scene.solve(w.estimator, w.pde, w.settings, *host_sets[s])
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0

    # ---- reduce over ranks: max time, summed work -----------------------------------------------
    if dist is not None:
        t = torch.tensor([dev_ms, e2e_s, walk_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_s, walk_ms = [float(v) for v in t.tolist()]
        c = torch.tensor([steps_dev, steps_e2e, prof_steps], dtype=torch.int64, device=dev)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        steps_dev, steps_e2e, prof_steps_all = [int(v) for v in c.tolist()]
    else:
        prof_steps_all = prof_steps

    # ---- parity smoke on the estimates of the last timed batch (exact solution is known) ----------
    hp, hs = host_sets[-1]
    rmse = float(np.sqrt(np.mean((hs.solutionMean - w.exact(hp.pt))**2)))

    if rank == 0:
        peak, peak_kind = measured_peaks()
        bytes_per_step = w.algorithmic_bytes_per_step()
        # per-rank kernel: this rank's steps / this rank's kernel time
        achieved = prof_steps*bytes_per_step/(local_walk_ms*1e-3)/1e9 if local_walk_ms > 0 else 0.0
        value = steps_dev/(dev_ms*1e-3)
        line = {
            "metric": "walk steps/sec", "value": value, "unit": "steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms/args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": w.name, "points_per_step_per_gpu": batch, "batch_pilot": pilot, "walks_per_point": w.n_walks,
                       "sharding": "evaluation points interleaved over ranks, BVH replicated, no data-path collective",
                       "l2": "inputs larger than L2: BVH %.0f MB + %.0f MB of per-walk records per step"
                             % ((scene.stats(0)["bytes"] + scene.stats(1)["bytes"])/1e6, walk_tasks/max(1, args.steps)*40/1e6),
                       "bvh_build_s": build_s, "rmse_vs_exact_last_batch": rmse},
            "e2e": {"value": steps_e2e/e2e_s, "unit": "steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1e3*e2e_s/max(1, e2e_steps), "steps": e2e_steps},
            "gpu_launches": int(launches_timed),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved/peak,
                         "traffic": None, "peak_kind": peak_kind, "kernel": "walkKernel<3,true>", "grid": grid,
                         "algorithmic_bytes_per_walk_step": bytes_per_step, "kernel_ms_per_step": walk_ms/args.steps,
                         "kernel_share_of_step": walk_ms/dev_ms if dev_ms > 0 else None,
                         "traffic_note": "ncu --set full of the round-1 kernel on a 512-point launch (profiles/README.md): 20 KB of "
                                         "DRAM traffic and 44 K warp instructions per walk step, 54 % of stalls on instruction "
                                         "fetch; issue/latency bound, not HBM bound (kernel reshaped since, not re-profiled)"},
        }
        if not args.no_cpu and world == 1:   # the CPU baseline is reported at N = 1 only
            line["cpu_baseline"] = cpu_baseline(args, w)
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def cpu_baseline(args, w):
    """The reference's TBB CPU solver (oracle/_ref) on a bounded random sample of the same workload."""
    from oracle import ref
    if not ref.available():
        return {"value": None, "unit": "steps/s", "cores": 0, "kind": "reference", "sample": "oracle/_ref not built"}
    scene = ref.RefScene(3, **w.scene_kwargs())
    n = args.ref_points or ref_sample_size(scene, w, 20.0, 16384)
    idx = np.sort(np.random.RandomState(123).choice(len(w.points), n, replace=False))
    steps, secs = ref_solve_sample(scene, w, idx)
    return {"value": steps/secs, "unit": "steps/s", "cores": ref.hardware_threads(), "kind": "reference",
            "sample": "%d random evaluation points x %d walks of the workload (%.1f s), TBB default arena, -O3 -DNDEBUG" % (n, w.n_walks, secs)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--batch", type=int, default=0, help="evaluation points per step per GPU (0: sized by a pilot solve)")
    ap.add_argument("--step-seconds", type=float, default=8.0, help="target duration of one step when --batch is 0")
    ap.add_argument("--subdiv", type=int, default=8, help="icosphere subdivisions (8 -> 1,310,720 triangles)")
    ap.add_argument("--points", type=int, default=1 << 20)
    ap.add_argument("--walks", type=int, default=64)
    ap.add_argument("--ref-points", type=int, default=0, help="evaluation points per CPU-reference sample (0: sized by a pilot run to about 20 s)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_native(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
