"""Aggregates an `ncu --page source --csv --print-source cuda,sass` export by source file and by line range:
samples, stall_no_inst share, executed warp instructions, lane efficiency. Usage: python scripts/ncu_by_function.py src_both.csv"""
import csv, sys, collections

rows = list(csv.reader(open(sys.argv[1])))
hdr = None
cur_file = None
per_file = collections.defaultdict(lambda: collections.defaultdict(float))
per_line = collections.defaultdict(lambda: collections.defaultdict(float))
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]; continue
    if len(r) >= 2 and r[0] == "Line No":
        hdr = r; continue
    if hdr is None or len(r) != len(hdr):
        continue
    if r[2] != "-":   # SASS row under a source line; the source-line row carries the aggregate
        continue
    d = dict(zip(hdr, r))
    def f(k):
        try: return float(d.get(k, 0) or 0)
        except ValueError: return 0.0
    key = (cur_file, int(r[0]))
    for k in ("# Samples", "Instructions Executed", "Thread Instructions Executed", "stall_no_inst", "stall_long_sb", "stall_wait",
              "stall_short_sb", "stall_branch_resolving", "stall_selected", "stall_not_selected", "stall_lg", "stall_mio", "stall_math", "stall_dispatch", "stall_barrier"):
        per_file[cur_file][k] += f(k)
        per_line[key][k] += f(k)
tot = sum(v["# Samples"] for v in per_file.values())
toti = sum(v["Instructions Executed"] for v in per_file.values())
print("total samples %d, warp instructions %.3g" % (tot, toti))
allst = collections.defaultdict(float)
for v in per_file.values():
    for k, x in v.items():
        if k.startswith("stall_"): allst[k] += x
print("stall shares:", ", ".join("%s %.1f%%" % (k[6:], 100*x/tot) for k, x in sorted(allst.items(), key=lambda t: -t[1])[:8]))
print("%-16s %8s %8s %8s %8s" % ("file", "samples%", "instr%", "no_inst%", "lanes"))
for fn, v in sorted(per_file.items(), key=lambda t: -t[1]["# Samples"]):
    print("%-16s %8.1f %8.1f %8.1f %8.1f" % (fn, 100*v["# Samples"]/tot, 100*v["Instructions Executed"]/toti, 100*v["stall_no_inst"]/max(1, v["# Samples"]),
          v["Thread Instructions Executed"]/max(1, v["Instructions Executed"])))
print("top lines:")
for (fn, ln), v in sorted(per_line.items(), key=lambda t: -t[1]["# Samples"])[:int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print("  %-14s:%-5d samples %5.2f%% instr %5.2f%% no_inst %4.0f%% long_sb %4.0f%% lanes %4.1f" % (fn, ln, 100*v["# Samples"]/tot, 100*v["Instructions Executed"]/toti,
          100*v["stall_no_inst"]/max(1, v["# Samples"]), 100*v["stall_long_sb"]/max(1, v["# Samples"]), v["Thread Instructions Executed"]/max(1, v["Instructions Executed"])))
