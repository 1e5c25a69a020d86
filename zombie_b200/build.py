"""Builds zombie_b200/libzombie_b200.so (sm_100a only) with nvcc. The library is built in-tree so
that it travels with the repository snapshot; there is no JIT and no CPU fallback."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libzombie_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-O3"]


def _sources():
    cu = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(".cu")]
    cpp = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(".cpp")]
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    inc = os.path.join(HERE, "..", "include")
    deps += [os.path.join(inc, "zombie_b200.h"), os.path.join(inc, "zombie_b200", "analytic.h")]
    return cu, cpp, deps


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    _, _, deps = _sources()
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, extra=()):
    if not force and not needs_build():
        return LIB
    cu, cpp, _ = _sources()
    cmd = [NVCC] + ARCH + COMMON + list(extra) + ["-shared", "-o", LIB] + cu + cpp
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd))
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed building libzombie_b200.so")
    return LIB


if __name__ == "__main__":
    build(force=True, verbose="-v" in sys.argv)
    print(LIB)
