#!/usr/bin/env python
"""Turns an .ncu-rep into a small text summary for profiles/: per kernel the headline counters (duration, registers,
occupancy limits, achieved warps, IPC, DRAM bytes, L1/L2 hit rates, thread efficiency), the warp-stall breakdown,
and — when the report carries source counters and the code was built with -lineinfo — the source lines that execute
the most instructions.   python scripts/summarize_ncu.py <report.ncu-rep> [--top 25] > profiles/<name>.txt"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "smsp__inst_executed_op_shared_atom.sum",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active", "sm__cycles_elapsed.avg"]


def ncu(args):
    return subprocess.run(["ncu"] + args, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout


def main():
    rep = sys.argv[1]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 25
    rows = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units = rows[0], rows[1]
    print("# ncu summary of %s" % rep)
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        print("\n## %s" % name[:150])
        for k in KEYS:
            if k in hdr:
                print("  %-72s %s %s" % (k, r[hdr.index(k)], units[hdr.index(k)]))
        stalls = []
        for i, h in enumerate(hdr):
            if "warp_issue_stalled" in h and h.endswith("per_warp_active.pct") and "not_issued" not in h:
                try:
                    stalls.append((float(r[i]), h.replace("smsp__warp_issue_stalled_", "").replace("_per_warp_active.pct", "")))
                except ValueError:
                    pass
            elif "warps_issue_stalled" in h and h.endswith("per_issue_active.ratio") and "not_issued" not in h:
                try:
                    stalls.append((float(r[i]), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
                except ValueError:
                    pass
        if stalls:
            tot = sum(v for v, _ in stalls) or 1.0
            print("  warp stall reasons (share of all stalled-warp samples): " + ", ".join("%s %.1f %%" % (n, 100*v/tot) for v, n in sorted(stalls, reverse=True)[:8]))
    src = ncu(["-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"])
    cur, hdr2, lines = None, None, []
    kernel = None
    for r in csv.reader(io.StringIO(src)):
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif r[0] == "Function Name":
            kernel = r[1]
        elif r[0] == "Line No":
            hdr2 = r
        elif hdr2 and r[0].isdigit() and "Instructions Executed" in hdr2:
            try:
                lines.append((int(r[hdr2.index("Instructions Executed")]), int(r[hdr2.index("# Samples")]), cur, int(r[0]), r[1].strip()[:100]))
            except (ValueError, IndexError):
                pass
    if lines:
        tot = float(sum(x[0] for x in lines)) or 1.0
        tots = float(sum(x[1] for x in lines)) or 1.0
        print("\n## instructions executed by source file (last kernel in the report: %s)" % (kernel or "?")[:100])
        by_file = defaultdict(lambda: [0, 0])
        for x in lines:
            by_file[x[2]][0] += x[0]
            by_file[x[2]][1] += x[1]
        for k, v in sorted(by_file.items(), key=lambda kv: -kv[1][0])[:10]:
            print("  %-28s inst %5.1f %%   stall samples %5.1f %%" % (k, 100*v[0]/tot, 100*v[1]/tots))
        print("\n## top %d source lines by instructions executed" % top)
        for x in sorted(lines, key=lambda x: -x[0])[:top]:
            print("  %5.2f %% inst %5.2f %% samples  %s:%d  %s" % (100*x[0]/tot, 100*x[1]/tots, x[2], x[3], x[4]))


if __name__ == "__main__":
    main()
