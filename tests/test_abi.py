"""CPU tests of the drop-in boundary: libzombie_b200.so loads, exports every symbol include/zombie_b200.h
declares, its struct layouts match the ctypes mirrors, and — with no CUDA device — it fails loudly instead of
falling back to a CPU path."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

from tests.conftest import ROOT
from zombie_b200 import abi, native

HEADER = os.path.join(ROOT, "include", "zombie_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(zb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = native.lib()
    names = declared_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), "libzombie_b200.so does not export %s" % name
    # and the python binding knows about all of them
    assert sorted(native.SYMBOLS) == names


def test_abi_version():
    assert native.lib().zb_abi_version() == 3   # 2: zb_pde_desc carries zb_grid_view members; 3: device builder, moving geometry, value channels


def test_struct_layouts_match_header(tmp_path):
    """sizeof/offsetof from a C compile of the header == the ctypes mirrors in zombie_b200/abi.py"""
    src = tmp_path/"layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "zombie_b200.h"\n'
                   'int main(void){printf("%zu %zu %zu %zu %zu %zu %zu\\n", sizeof(zb_walk_settings), sizeof(zb_pde_desc), '
                   'sizeof(zb_sample_points), sizeof(zb_statistics), offsetof(zb_pde_desc, robinCoeff), '
                   'offsetof(zb_pde_desc, kelvinOrigin), offsetof(zb_sample_points, rngState));return 0;}\n')
    exe = tmp_path/"layout"
    subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    want = [C.sizeof(abi.WalkSettings), C.sizeof(abi.PdeDesc), C.sizeof(abi.SamplePoints), C.sizeof(abi.Statistics),
            abi.PdeDesc.robinCoeff.offset, abi.PdeDesc.kelvinOrigin.offset, abi.SamplePoints.rngState.offset]
    assert got == want


def test_header_is_plain_c(tmp_path):
    src = tmp_path/"c89.c"
    src.write_text('#include "zombie_b200.h"\nint main(void){return ZB_OK;}\n')
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), "-c", str(src),
                           "-o", str(tmp_path/"c89.o")])


def test_no_cpu_fallback_without_device():
    """on a box without a GPU the library must refuse to create a context (there is no CPU execution path)"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    assert native.device_count() == 0
    with pytest.raises(native.ZombieError) as e:
        native.Context(0)
    assert "no CUDA device" in str(e.value)


def test_invalid_arguments_are_reported_not_fatal():
    lib = native.lib()
    assert lib.zb_context_create(0, None) == 1  # ZB_ERR_INVALID_ARGUMENT
    assert b"out is NULL" in lib.zb_last_error()
    assert lib.zb_solve(None, 0, None, None, None, None, None) == 1
    assert lib.zb_closest_point(None, 0, 1, None, None, None, None, None, None) == 1


def test_product_never_imports_the_oracle():
    """only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference legs may touch oracle/"""
    pkg = os.path.join(ROOT, "zombie_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                text = open(os.path.join(base, f), errors="ignore").read()
                assert "oracle" not in text.replace("test oracle", "").replace("the oracle's wrapper", ""), os.path.join(base, f)
    for f in os.listdir(os.path.join(ROOT, "include")):
        pass
