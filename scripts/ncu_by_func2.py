"""Aggregates an `ncu --page source --csv --print-source cuda,sass` export by the FUNCTION a source line belongs to
(function = nearest preceding definition line in the .cuh file): share of executed warp instructions, of stall samples,
lane efficiency. Usage: python scripts/ncu_by_func2.py src.csv [csrc dir]"""
import csv, sys, collections, re, os, bisect
csrc = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "zombie_b200", "csrc")
defs = {}
pat = re.compile(r'^\s*(?:template\s*<[^>]*>\s*)?(?:ZB_DEV_COLD|ZB_DEV|ZB_COOP_NOINLINE|ZB_COOP|ZB_HOSTDEV|__global__|inline)\b[^;]*?([A-Za-z_][A-Za-z0-9_]*)\s*(?:<[^>]*>)?\s*\(')
for fn in os.listdir(csrc):
    if not fn.endswith((".cuh", ".h", ".cu")): continue
    L = []
    prev_template = False
    for i, line in enumerate(open(os.path.join(csrc, fn)), 1):
        m = pat.match(line)
        if m and not line.strip().startswith("//") and not line.rstrip().endswith(";"):
            L.append((i, m.group(1)))
    defs[fn] = L
def func_of(fn, ln):
    L = defs.get(fn)
    if not L: return fn
    k = bisect.bisect_right([a for a, _ in L], ln) - 1
    return "%s:%s" % (fn, L[k][1] if k >= 0 else "?")
rows = csv.reader(open(sys.argv[1]))
hdr = None; cur = None
agg = collections.defaultdict(lambda: collections.defaultdict(float))
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if len(r) >= 2 and r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) != len(hdr) or r[2] != "-": continue
    d = dict(zip(hdr, r))
    f = func_of(cur, int(r[0]))
    for k in ("# Samples", "Instructions Executed", "Thread Instructions Executed", "stall_long_sb", "stall_no_inst", "stall_wait", "stall_short_sb", "stall_branch_resolving"):
        try: agg[f][k] += float(d.get(k, 0) or 0)
        except ValueError: pass
ts = sum(v["# Samples"] for v in agg.values()); ti = sum(v["Instructions Executed"] for v in agg.values())
print("%-44s %8s %8s %6s %8s %8s" % ("function", "samples%", "instr%", "lanes", "long_sb%", "no_inst%"))
for f, v in sorted(agg.items(), key=lambda t: -t[1]["# Samples"])[:45]:
    print("%-44s %8.1f %8.1f %6.1f %8.0f %8.0f" % (f, 100*v["# Samples"]/ts, 100*v["Instructions Executed"]/ti, v["Thread Instructions Executed"]/max(1, v["Instructions Executed"]),
          100*v["stall_long_sb"]/max(1, v["# Samples"]), 100*v["stall_no_inst"]/max(1, v["# Samples"])))
