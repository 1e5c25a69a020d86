"""profiles/r2_ncu_evidence.json from an ncu summary (scripts/summarize_ncu.py output) of one steady-state round of the wavefront:
DRAM bytes read + written per walk step summed over the round's kernels, and the share / issue statistics of the dominant one.
    python scripts/ncu_evidence.py profiles/r2_wave_round_ncu.txt
A round advances every live walk by one step; in the captured round every slot holds a walk, so the walk steps of the round are the
positions the step kernel covers (grid x 128)."""
import json, os, re, sys

path = sys.argv[1]
kernels, cur = {}, None
for line in open(path):
    m = re.match(r"## void (\w+)<", line)
    if m:
        cur = kernels.setdefault(m.group(1), {})
        continue
    if cur is None or line.startswith("## "):
        if line.startswith("## ") and not m:
            cur = None
        continue
    t = line.split()
    if len(t) >= 2 and re.match(r"^[a-z_]+__", t[0]):
        try:
            v = float(t[1])
        except ValueError:
            continue
        unit = t[2] if len(t) > 2 else ""
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1.0, "us": 1e-3, "s": 1e3}.get(unit, 1.0)
        cur[t[0]] = v*scale
steps = kernels["waveStepKernel"]["launch__grid_size"]*128
total_bytes = sum(k.get("dram__bytes_read.sum", 0.0) + k.get("dram__bytes_write.sum", 0.0) for k in kernels.values())
total_ms = sum(k["gpu__time_duration.sum"] for k in kernels.values())
dom = max(kernels, key=lambda n: kernels[n]["gpu__time_duration.sum"])
out = {
    "profile": os.path.relpath(path, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")),
    "dram_bytes_per_walk_step": total_bytes/steps,
    "walk_steps_in_captured_round": int(steps),
    "round_ms_under_ncu": total_ms,
    "kernel_ms_under_ncu": {n: k["gpu__time_duration.sum"] for n, k in kernels.items()},
    "dominant_kernel": dom + "<3>",
    "dominant_kernel_share_of_round": kernels[dom]["gpu__time_duration.sum"]/total_ms,
    "dominant_kernel_issue_active_pct": kernels[dom].get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
    "dominant_kernel_threads_per_instruction": kernels[dom].get("smsp__thread_inst_executed_per_inst_executed.ratio"),
}
dst = os.path.join(os.path.dirname(os.path.abspath(path)), "r2_ncu_evidence.json")
json.dump(out, open(dst, "w"), indent=1)
print(json.dumps(out, indent=1))
