"""Builds zombie_b200/libzombie_b200.so (sm_100a only) with nvcc. The library is built in-tree so
that it travels with the repository snapshot; there is no JIT and no CPU fallback. Translation units are
compiled in parallel into zombie_b200/build/*.o and linked; only stale objects are rebuilt."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libzombie_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-O3"]


def _sources():
    srcs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(".cu") or f.endswith(".cpp")]
    inc = os.path.join(HERE, "..", "include")
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh") or f.endswith(".h")]
    headers += [os.path.join(inc, "zombie_b200.h"), os.path.join(inc, "zombie_b200", "analytic.h"), os.path.abspath(__file__)]
    return srcs, headers


def _obj(src):
    return os.path.join(OBJ, os.path.basename(src) + ".o")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def needs_build():
    srcs, headers = _sources()
    return _stale(LIB, srcs + headers)


def build(force=False, verbose=False, extra=()):
    if not force and not needs_build():
        return LIB
    srcs, headers = _sources()
    os.makedirs(OBJ, exist_ok=True)

    def compile_one(src):
        obj = _obj(src)
        if not force and not _stale(obj, [src] + headers):
            return 0, ""
        cmd = [NVCC] + ARCH + COMMON + list(extra) + (["-Xptxas=-v"] if verbose else []) + ["-c", src, "-o", obj]
        if verbose:
            print(" ".join(cmd))
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        return r.returncode, r.stdout

    with ThreadPoolExecutor(max_workers=max(1, min(len(srcs), os.cpu_count() or 1))) as pool:
        results = list(pool.map(compile_one, srcs))
    failed = [out for rc, out in results if rc != 0]
    if verbose or failed:
        sys.stderr.write("".join(out for _, out in results))
    if failed:
        raise RuntimeError("nvcc failed building libzombie_b200.so")
    r = subprocess.run([NVCC] + ARCH + ["-shared", "-o", LIB] + [_obj(s) for s in srcs], stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("linking libzombie_b200.so failed")
    return LIB


def build_variant(name, extra):
    """A/B builds for GPU measurements: zombie_b200/variants/lib<name>.so compiled with extra nvcc flags
    (e.g. -DZB_UNROLL_CHILD_LOOPS); select one at run time with ZOMBIE_B200_LIB (scripts/gpu_variants.sh)."""
    srcs, _ = _sources()
    out_dir = os.path.join(HERE, "variants")
    obj_dir = os.path.join(OBJ, "variant_" + name)
    os.makedirs(out_dir, exist_ok=True)
    os.makedirs(obj_dir, exist_ok=True)

    def compile_one(src):
        obj = os.path.join(obj_dir, os.path.basename(src) + ".o")
        r = subprocess.run([NVCC] + ARCH + COMMON + list(extra) + ["-c", src, "-o", obj], stdout=subprocess.PIPE,
                           stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise RuntimeError(r.stdout)
        return obj

    with ThreadPoolExecutor(max_workers=max(1, min(len(srcs), os.cpu_count() or 1))) as pool:
        objs = list(pool.map(compile_one, srcs))
    lib = os.path.join(out_dir, "lib%s.so" % name)
    r = subprocess.run([NVCC] + ARCH + ["-shared", "-o", lib] + objs, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(r.stdout)
    return lib


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--variant":  # python -m zombie_b200.build --variant NAME [nvcc flags...]
        print(build_variant(sys.argv[2], sys.argv[3:]))
    else:
        build(force="-f" in sys.argv, verbose="-v" in sys.argv)
        print(LIB)
