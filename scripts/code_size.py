#!/usr/bin/env python3
"""Static SASS size of every kernel in the built library and, for one kernel, the instruction count per inlined call
site (needs -lineinfo, which zombie_b200/build.py always passes). The walk kernel is bound by instruction fetch
(profiles/README.md: 54 % of the warp-stall samples are stall_no_inst), so the size of the code a traversal loop
touches is a first-order performance figure that can be read without a GPU.

    python scripts/code_size.py                        # all kernels
    python scripts/code_size.py walkKernelILi3ELb1     # + per-call-site breakdown of the matching kernel
"""
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "zombie_b200", "build", "capi.cu.o")
INST = re.compile(r"^\s+/\*[0-9a-f]+\*/")
ANN = re.compile(r'//## File "([^"]+)", line (\d+)')


def kernel_sizes():
    out = subprocess.run(["cuobjdump", "-sass", OBJ], capture_output=True, text=True, check=True).stdout
    sizes, name = collections.Counter(), None
    for line in out.split("\n"):
        if "Function :" in line:
            name = line.split()[-1]
        elif INST.match(line) and name:
            sizes[name] += 1
    return sizes


def demangle(names):
    try:
        out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True, check=True).stdout
        return dict(zip(names, out.split("\n")))
    except Exception:
        return {n: n for n in names}


def call_sites(pattern):
    with tempfile.TemporaryDirectory() as tmp:
        subprocess.run(["cuobjdump", "-xelf", "all", OBJ], cwd=tmp, check=True, capture_output=True)
        cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
        dis = subprocess.run(["nvdisasm", "-gi", os.path.join(tmp, cubin)], capture_output=True, text=True, check=True).stdout
    on, pending, cur = False, [], None
    outer, inner = collections.Counter(), collections.Counter()
    for line in dis.split("\n"):
        if line.startswith("//--------------------- .text."):
            on = pattern in line
            continue
        if not on:
            continue
        m = ANN.search(line)
        if m:
            pending.append((os.path.basename(m.group(1)), int(m.group(2))))
            continue
        if INST.match(line):
            if pending:
                cur, pending = pending, []
            if cur:
                outer[cur[-1]] += 1
                inner[(cur[-1], cur[-2] if len(cur) > 1 else ("-", 0))] += 1
    return outer, inner


def main():
    sizes = kernel_sizes()
    names = demangle(list(sizes))
    print("SASS instructions (16 B each) per kernel, device functions it calls included:")
    for n, c in sizes.most_common():
        short = re.sub(r"\(anonymous namespace\)::|zb::", "", names[n]).split("(")[0]
        print("  %6d  %6.1f KB  %s" % (c, c*16/1024, short))
    if len(sys.argv) > 1:
        outer, inner = call_sites(sys.argv[1])
        total = sum(outer.values())
        print("\n%s: %d instructions by outermost source line" % (sys.argv[1], total))
        for (f, l), c in outer.most_common(12):
            print("  %6d  %5.1f %%  %s:%d" % (c, 100.0*c/total, f, l))
        print("by (outermost line <- first inlined frame):")
        for ((f, l), (g, k)), c in inner.most_common(24):
            print("  %6d  %5.1f %%  %s:%d <- %s:%d" % (c, 100.0*c/total, f, l, g, k))


if __name__ == "__main__":
    main()
