#!/usr/bin/env python
"""Benchmark of the walk hot path: BASELINE.json's metric "walk steps/sec" on the headline workload C3
(3D screened Poisson, Robin + Dirichlet boundary of a 1,310,720-triangle bumpy icosphere, WalkOnStars
with gradient estimates, 1,048,576 evaluation points x 64 walks).

A *step* is one zb_solve over one batch of --batch evaluation points of that workload per GPU (a different,
contiguous slice of the rank's points each step; default 32,768 points = 2,097,152 walks per GPU per step).
`value` is whole-job walk steps/s with the points already resident in HBM (zb_solve_dev); `e2e` is the same
through the host-buffer C-ABI call zb_solve with the host<->device copies of every in/out array inside the
timed region.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--scaling weak|strong]

N > 1: launched under torchrun, one rank per GPU; the evaluation points are sharded over the ranks (interleaved;
the BVH is replicated) and every step ends with ONE collective inside the timed region, the NCCL all-gather of the
per-point statistics (56 B/point) from the device buffers the solve wrote. Timing is the max over ranks; the
spread of the per-rank kernel time (load imbalance) and the gather time are reported in `multi_gpu`.
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


def ref_sample_size(scene, w, target_s, cores, cap=16384):
    """Evaluation points the reference solves in about target_s seconds on this host (pilot run), but at least 64 per
    hardware thread where the time allows at least 24: walk lengths are geometric-tailed, and a handful of points per thread
    leaves TBB's workers idle at the end of a solve."""
    rng = np.random.RandomState(321)
    n0 = 4*cores
    _, secs = ref_solve_sample(scene, w, np.sort(rng.choice(len(w.points), n0, replace=False)))
    n = int(n0*target_s/max(secs, 1e-3))
    return int(min(cap, max(24*cores, min(n, 64*cores) if n < 64*cores else n)))


def ref_builds():
    """the reference builds this host can run: the parity build, and the -march=x86-64-v3 one (BASELINE.md §2 allows the
    timed arm the vector ISA of a native Release build)"""
    from oracle import ref
    builds = [("", "-O3 -DNDEBUG")]
    if ref.cpu_supports_v3() and os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libzombie_ref_v3.so")):
        builds.append(("v3", "-O3 -DNDEBUG -march=x86-64-v3"))
    return builds


def run_reference(args, rank, world):
    if rank != 0:
        return
    from oracle import ref
    if not ref.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libzombie_ref.so has not been built"}))
        return
    w = workloads.config3(args.subdiv, args.points, args.walks)
    cores = ref.hardware_threads()
    # pick the faster of the two builds on a short pilot, then time it
    pilot = {}
    idx0 = np.sort(np.random.RandomState(77).choice(len(w.points), 8*cores, replace=False))
    scenes = {}
    for name, flags in ref_builds():
        ref.select_build(name)
        scenes[name] = ref.RefScene(3, **w.scene_kwargs())
        steps, secs = ref_solve_sample(scenes[name], w, idx0)
        pilot[flags] = steps/secs
    best = max(ref_builds(), key=lambda b: pilot[b[1]])
    scene = scenes[best[0]]
    n_runs = args.warmup + args.steps
    n = args.ref_points or ref_sample_size(scene, w, min(20.0, 240.0/max(1, n_runs)), cores)
    rng = np.random.RandomState(123)
    total_steps, total_secs = 0, 0.0
    for s in range(n_runs):
        idx = np.sort(rng.choice(len(w.points), n, replace=False))
        steps, secs = ref_solve_sample(scene, w, idx)
        if s >= args.warmup:
            total_steps += steps
            total_secs += secs
    value = total_steps/total_secs
    sample = ("%d random evaluation points (%.0f per thread) x %d walks of the workload per step, TBB default arena (%d threads), %s"
              % (n, n/float(cores), w.n_walks, cores, best[1]))
    print(json.dumps({
        "impl": "reference", "metric": "walk steps/sec", "value": value, "unit": "steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3*total_secs/args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": w.name},
        "cpu_baseline": {"value": value, "unit": "steps/s", "cores": cores, "kind": "reference", "sample": sample,
                         "builds_pilot_steps_per_s": pilot},
        "e2e": {"value": value, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


# ------------------------------------------------------------------------------------------------
# native arm
# ------------------------------------------------------------------------------------------------
POINT_KEYS = ("pt", "normal", "directionForDerivative", "type", "estimationQuantity", "distToAbsorbingBoundary",
              "distToReflectingBoundary", "estimateBoundaryNormalAligned", "nWalks", "rngState", "rngInc", "firstSphereRadius")
STAT_KEYS = ("solutionMean", "solutionM2", "gradientMean", "gradientM2", "totalFirstSourceContribution",
             "totalDerivativeContribution", "nSolutionEstimates", "nGradientEstimates", "totalWalkLength", "totalSplits")

# ncu evidence of the shipped kernels: profiles/r2_ncu_evidence.json, written by scripts/ncu_evidence.py from the --set full
# capture of one steady-state round of this workload (profiles/r2_wave_round_ncu.txt: 984,320 live walks) — DRAM bytes read +
# written per walk step summed over the round's kernels, and the issue / lane statistics of the dominant kernel
NCU_EVIDENCE = {
    "profile": "profiles/r2_wave_round_ncu.txt",
    "dram_bytes_per_walk_step": None,
    "dominant_kernel": "waveStarPooledKernel<3>",
}


def load_ncu_evidence():
    path = os.path.join(ROOT, "profiles", "r2_ncu_evidence.json")
    if os.path.exists(path):
        with open(path) as f:
            NCU_EVIDENCE.update(json.load(f))
    return NCU_EVIDENCE


def run_native(args, rank, world, local_rank):
    import torch
    from zombie_b200 import native, sharding

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
    # the same hierarchies from the on-device builder (zb_context_set_bvh_builder; setup, untimed): built twice, the second
    # build reuses the builder's scratch — what a rebuild of moving geometry costs
    device_build_ms = device_rebuild_ms = None
    if rank == 0:
        kw = w.scene_kwargs()
        dscene = native.Scene(ctx, 3, builder=abi.BVH_BUILDER_DEVICE, **kw)
        device_build_ms = sum(dscene.build_profile(b)["ms"] for b in (abi.ABSORBING, abi.REFLECTING))   # a fresh scene: includes cudaMalloc of its arrays
        dscene.update_vertices(abi.ABSORBING, kw["absorbing"][0])
        dscene.update_vertices(abi.REFLECTING, kw["reflecting"][0])
        device_rebuild_ms = sum(dscene.build_profile(b)["ms"] for b in (abi.ABSORBING, abi.REFLECTING))  # what moving geometry pays per update
        dscene.close()

    # this rank's shard of the evaluation points (interleaved for load balance); distances to the two
    # boundaries are part of SamplePoint construction in the reference (demo.cpp:9-46) -> setup, untimed
    shard = sharding.shard_indices(len(w.points), rank, world)
    n_steps = args.warmup + args.steps
    _, dist_a, _, _ = scene.closest_point(abi.ABSORBING, w.points, want_normal=False)
    _, dist_r, _, _ = scene.closest_point(abi.REFLECTING, w.points, want_normal=False)

    def to_dev(a):
        return torch.from_numpy(np.ascontiguousarray(a)).to(dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- batch: the evaluation points one step solves per GPU ("weak": fixed per GPU; "strong": fixed in total)
    batch = args.batch if args.scaling == "weak" else max(1, args.batch//world)
    batch = min(batch, len(shard))

    def batch_indices(s):
        start = (s*batch) % max(1, len(shard) - batch + 1)
        return shard[start:start + batch]

    def device_set(idx, seed_base):
        hp = make_points(w, idx, dist_a, dist_r, seed_base)
        tensors = {}
        for k in POINT_KEYS:
            a = getattr(hp, k)
            tensors[k] = to_dev(a.view(np.int64) if a.dtype == np.uint64 else a)
        stats = sharding.DeviceStatistics(hp.n, 3, dev)          # one contiguous buffer: the gather is one collective
        tensors["steps"] = torch.zeros(1, dtype=torch.int64, device=dev)
        ps = abi.SamplePoints(hp.n, *[tensors[k].data_ptr() for k in POINT_KEYS])
        return ps, stats.as_struct(), tensors, stats

    gathered = torch.empty((world, (2 + 2*3 + 2 + 4)*batch), dtype=torch.float32, device=dev) if dist is not None else None

    # ---- device-resident arm --------------------------------------------------------------------
    dev_steps = [device_set(batch_indices(s), 0) for s in range(n_steps)]
    gather_ev = []

    def run_dev(s, timed=False):
        ps, ss, tensors, stats = dev_steps[s]
        scene.solve_dev(w.estimator, w.pde, w.settings, ps, ss, tensors["steps"].data_ptr())
        if dist is not None:
            # the path's one collective: every rank ends with the statistics of all ranks' points (56 B/point), gathered
            # from the device buffers zb_solve_dev wrote, NCCL over NVLink
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            stats.all_gather(gathered)
            e1.record(stream)
            if timed:
                gather_ev.append((e0, e1))

    for s in range(args.warmup):
        run_dev(s)
    barrier()
    launches_before = ctx.last_solve_profile()["launches"] if args.warmup > 0 else 0
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    walk_ms, walk_tasks, rounds = 0.0, 0, 0
    ev0.record(stream)
    for s in range(args.warmup, n_steps):
        run_dev(s, timed=True)
        # the library brackets the walk phase of every solve with CUDA events on the launching stream
        prof = ctx.last_solve_profile()
        walk_ms += prof["walk_ms"]
        walk_tasks += prof["walks"]
    ev1.record(stream)
    barrier()
    dev_ms = ev0.elapsed_time(ev1)
    clocks = sampler.summary()
    launches_timed = ctx.last_solve_profile()["launches"] - launches_before
    steps_dev = int(sum(int(dev_steps[s][2]["steps"].item()) for s in range(args.warmup, n_steps)))
    gather_ms = float(sum(a.elapsed_time(b) for a, b in gather_ev))
    prof_steps, local_walk_ms = steps_dev, walk_ms  # this rank's kernels: its steps / its kernel time
    rmse_dev = float(np.sqrt(np.mean((dev_steps[-1][3].to_host().solutionMean - w.exact(w.points[batch_indices(n_steps - 1)]))**2)))

    # ---- end-to-end arm --------------------------------------------------------------------------
    # N = 1: host buffers through zb_solve, the reference-facing call (copies inside the library, inside the timed region).
    # N > 1: the same per rank, written out — pinned host inputs -> device, zb_solve_dev, all-gather of the statistics on the
    # device, gathered statistics -> pinned host on every rank — because the gather has to sit between the solve and the
    # read-back. (At most 3 timed steps after 1 warm-up: this arm repeats the device arm's work plus the copies.)
    e2e_warmup, e2e_steps = min(1, args.warmup), min(3, args.steps)
    host_sets = []
    for s in range(e2e_warmup + e2e_steps):
        hp = make_points(w, batch_indices(s), dist_a, dist_r, 1 << 40)
        hs = abi.HostStatistics(hp.n, 3)
        for obj, keys in ((hp, POINT_KEYS), (hs, STAT_KEYS)):   # pinned host memory
            for k in keys:
                a = getattr(obj, k)
                v = a.view(np.int64) if a.dtype == np.uint64 else a
                t = torch.from_numpy(v.copy()).pin_memory()
                pinned = t.numpy()
                setattr(obj, k, pinned.view(np.uint64) if a.dtype == np.uint64 else pinned)
                obj.__dict__.setdefault("_keep", {})[k] = t
        host_sets.append((hp, hs))
    n, dim = host_sets[0][0].n, 3
    in_bytes = n*(4*dim*3 + 4*6 + 1 + 8*2)
    stat_bytes = n*(4*2 + 4*dim*2 + 4*2 + 4*4)
    if dist is None:
        h2d, d2h = in_bytes + stat_bytes, n*(8 + 4) + stat_bytes + 8   # zb_solve: points + statistics in; rng state, radius, statistics, step count out
        e2e_path = "zb_solve on pinned host buffers (host<->device copies inside the call)"

        def run_e2e(s):
            return scene.solve(w.estimator, w.pde, w.settings, *host_sets[s])
    else:
        h2d, d2h = in_bytes, world*stat_bytes + 8
        e2e_path = "pinned host inputs -> device, zb_solve_dev, NCCL all-gather of the statistics, all ranks' statistics -> pinned host"
        dpts = {k: torch.empty_like(host_sets[0][0]._keep[k], device=dev) for k in POINT_KEYS}
        dstats = sharding.DeviceStatistics(n, 3, dev)
        dsteps = torch.zeros(1, dtype=torch.int64, device=dev)
        host_gathered = torch.empty((world, dstats.words*n), dtype=torch.float32).pin_memory()
        host_count = torch.zeros(1, dtype=torch.int64).pin_memory()
        ps_e2e = abi.SamplePoints(n, *[dpts[k].data_ptr() for k in POINT_KEYS])

        def run_e2e(s):
            hp, _ = host_sets[s]
            for k in POINT_KEYS:
                dpts[k].copy_(hp._keep[k], non_blocking=True)
            dstats.buf.zero_()
            scene.solve_dev(w.estimator, w.pde, w.settings, ps_e2e, dstats.as_struct(), dsteps.data_ptr())
            dstats.all_gather(gathered)
            host_gathered.copy_(gathered, non_blocking=True)
            host_count.copy_(dsteps, non_blocking=True)
            torch.cuda.synchronize()
            return int(host_count.item())
    for s in range(e2e_warmup):
        run_e2e(s)
    barrier()
    t0 = time.perf_counter()
    steps_e2e = 0
    for s in range(e2e_warmup, e2e_warmup + e2e_steps):
        steps_e2e += run_e2e(s)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0

    # ---- reduce over ranks: max time, summed work, spread of the per-rank kernel time (load imbalance) ----
    walk_ms_min = walk_ms_max = walk_ms
    if dist is not None:
        t = torch.tensor([dev_ms, e2e_s, walk_ms, gather_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_s, walk_ms_max, gather_ms = [float(v) for v in t.tolist()]
        t = torch.tensor([walk_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        walk_ms_min = float(t.item())
        c = torch.tensor([steps_dev, steps_e2e], dtype=torch.int64, device=dev)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        steps_dev, steps_e2e = [int(v) for v in c.tolist()]

    if rank == 0:
        peak, peak_kind = measured_peaks()
        ev = load_ncu_evidence()
        bytes_per_step = w.algorithmic_bytes_per_step()
        # per-rank kernels: this rank's steps / this rank's kernel time
        achieved = prof_steps*bytes_per_step/(local_walk_ms*1e-3)/1e9 if local_walk_ms > 0 else 0.0
        value = steps_dev/(dev_ms*1e-3)
        traffic = ev.get("dram_bytes_per_walk_step")
        line = {
            "metric": "walk steps/sec", "value": value, "unit": "steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms/args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": w.name, "points_per_step_per_gpu": batch, "walks_per_point": w.n_walks,
                       "walks_per_step_per_gpu": walk_tasks//max(1, args.steps),
                       "sharding": "evaluation points interleaved over ranks, BVH replicated; one NCCL all-gather of the per-point "
                                   "statistics (56 B/point, device buffers) per step inside the timed region" if world > 1 else
                                   "single GPU",
                       "l2": "inputs larger than L2: BVH %.0f MB + %.0f MB of walk state and per-walk records per step"
                             % ((scene.stats(0)["bytes"] + scene.stats(1)["bytes"])/1e6, (262144*128 + walk_tasks/max(1, args.steps)*40)/1e6),
                       "bvh_build_s": build_s, "bvh_build_on_device_ms": device_build_ms, "bvh_rebuild_on_device_ms": device_rebuild_ms,
                       "rmse_vs_exact_last_batch": rmse_dev, "e2e_path": e2e_path},
            "e2e": {"value": steps_e2e/e2e_s, "unit": "steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1e3*e2e_s/max(1, e2e_steps), "steps": e2e_steps},
            "gpu_launches": int(launches_timed),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved/peak,
                         "traffic": None if traffic is None else traffic*prof_steps/max(1, args.steps),
                         "peak_kind": peak_kind,
                         "kernel": "one solve's wavefront rounds: waveClosestKernel + waveStarPrivateKernel + waveStarPooledKernel + waveStepKernel (+ cub sort), then waveTailKernel",
                         "per_launch": "one zb_solve_dev (= one step)",
                         "algorithmic_bytes_per_walk_step": bytes_per_step, "dram_bytes_per_walk_step_ncu": traffic,
                         "kernel_ms_per_step": local_walk_ms/args.steps,
                         "kernel_share_of_step": local_walk_ms/dev_ms if dev_ms > 0 else None,
                         "ncu": {k: v for k, v in ev.items() if k != "dram_bytes_per_walk_step"}},
        }
        if world > 1:
            # the gather's events bracket the collective on each rank, so its time includes waiting for the slowest rank's solve
            line["multi_gpu"] = {"gather_ms_per_step": gather_ms/args.steps, "gather_includes_wait_for_slowest_rank": True,
                                 "gather_bytes_per_step": world*stat_bytes,
                                 "walk_ms_per_step_min_rank": walk_ms_min/args.steps, "walk_ms_per_step_max_rank": walk_ms_max/args.steps}
        if not args.no_cpu and world == 1:   # the CPU baseline is reported at N = 1 only
            line["cpu_baseline"] = cpu_baseline(args, w)
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def cpu_baseline(args, w):
    """The reference's TBB CPU solver (oracle/_ref) on a bounded random sample of the same workload, both builds."""
    from oracle import ref
    if not ref.available():
        return {"value": None, "unit": "steps/s", "cores": 0, "kind": "reference", "sample": "oracle/_ref not built"}
    cores = ref.hardware_threads()
    out = {}
    n = None
    for name, flags in ref_builds():
        ref.select_build(name)
        scene = ref.RefScene(3, **w.scene_kwargs())
        if n is None:
            n = args.ref_points or ref_sample_size(scene, w, 12.0, cores)
        idx = np.sort(np.random.RandomState(123).choice(len(w.points), n, replace=False))
        steps, secs = ref_solve_sample(scene, w, idx)
        out[flags] = (steps/secs, secs)
        scene.close()
    best = max(out, key=lambda k: out[k][0])
    return {"value": out[best][0], "unit": "steps/s", "cores": cores, "kind": "reference",
            "sample": "%d random evaluation points (%.0f per thread) x %d walks of the workload (%.1f s), TBB default arena, %s"
                      % (n, n/float(cores), w.n_walks, out[best][1], best),
            "builds_steps_per_s": {k: v[0] for k, v in out.items()}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--batch", type=int, default=65536, help="evaluation points per step (per GPU when --scaling weak, in total when strong)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
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
