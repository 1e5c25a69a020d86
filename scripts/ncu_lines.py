"""Per-source-line view of ONE kernel of an ncu report: python scripts/ncu_lines.py <rep> <kernel regex> [bucket]
Lines are grouped in buckets of `bucket` source lines; prints share of executed instructions, stall samples and of the
no_instruction stall samples per bucket (where a kernel waits for instruction fetch)."""
import csv, io, subprocess, sys
from collections import defaultdict

rep, rx = sys.argv[1], sys.argv[2]
bucket = int(sys.argv[3]) if len(sys.argv) > 3 else 20
out = subprocess.run(["ncu", "-i", rep, "--kernel-name", "regex:" + rx, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
cur, hdr, g = None, None, defaultdict(lambda: [0, 0, 0, 0])
for r in csv.reader(io.StringIO(out)):
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = r
    elif hdr and r[0].isdigit() and len(r) > hdr.index("stall_no_inst"):
        try:
            k = (cur, int(r[0])//bucket*bucket)
            g[k][0] += int(r[hdr.index("Instructions Executed")])
            g[k][1] += int(r[hdr.index("# Samples")])
            g[k][2] += int(r[hdr.index("stall_no_inst")])
            g[k][3] += int(r[hdr.index("stall_long_sb")])
        except ValueError:
            pass
tot = [sum(v[i] for v in g.values()) or 1 for i in range(4)]
print("# %s  (%d executed warp instructions, %d stall samples)" % (rx, tot[0], tot[1]))
print("# file:first line of bucket      exec %   samples %   no_inst %   long_sb %")
for k, v in sorted(g.items(), key=lambda kv: -kv[1][1])[:45]:
    print("%-22s %6d   %6.1f %9.1f %10.1f %10.1f" % (k[0], k[1], 100*v[0]/tot[0], 100*v[1]/tot[1], 100*v[2]/tot[2], 100*v[3]/tot[3]))
