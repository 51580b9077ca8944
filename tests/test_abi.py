"""CPU-only: the C-ABI library loads, exports every symbol include/b200_sixdof.h
declares, and fails loudly (no CPU fallback) when there is no GPU."""

import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

import elodin_b200 as el
from elodin_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "b200_sixdof.h")


def _header_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def built_lib():
    if not os.path.exists(_lib.LIB_PATH):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "elodin_b200", "csrc")], check=True)
    return ctypes.CDLL(_lib.LIB_PATH)


def test_header_symbols_all_exported(built_lib):
    syms = _header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(built_lib, s), f"{s} declared in include/b200_sixdof.h but not exported"
    assert sorted(_lib.SYMBOLS) == syms  # the ctypes binding covers the whole header


def test_component_id_matches_reference_table(built_lib):
    """FNV-1a-64 & ~(1<<63): the ids SURVEY §8a-7 lists, host (pure Python) and library agree."""
    want = {"world_accel": 0x019091805BC057F4, "simulation_time_step": 0x08E7DDBB2CCEAAB5, "tick": 0x1E7683EF2EBC7684,
            "world_vel": 0x4B03B28A841EDD5F, "world_pos": 0x5D1C198A8E96E26E, "inertia": 0x5FD14829C04C0F91,
            "force": 0x675AD8AFB3EEEBE4}
    built_lib.b200_component_id.restype = ctypes.c_uint64
    built_lib.b200_component_id.argtypes = [ctypes.c_char_p]
    for name, cid in want.items():
        assert el.component_id(name) == cid
        assert built_lib.b200_component_id(name.encode()) == cid
    ordered = sorted(want, key=want.get)
    assert ordered == ["world_accel", "simulation_time_step", "tick", "world_vel", "world_pos", "inertia", "force"]
    # from_pair("a","b") == new("a.b") (types.rs:47-57)
    assert el.component_id("a.b") == built_lib.b200_component_id(b"a.b")


def test_struct_layouts_match_header(tmp_path):
    assert ctypes.sizeof(_lib.Effector) == 4 + 4 + 64 + 8 + 4 + 4 + 8 + 8 + 8 + 8 + 8 + 8 + 8  # ABI v3: + table0, table1, table_len
    assert ctypes.sizeof(_lib.Desc) == 16 + 16 + 16 + 8 + 4 + 4 + 4 + 4 + 8 + 4 + 4
    assert ctypes.sizeof(_lib.Timings) == 48
    # and against the C compiler's view of include/b200_sixdof.h: sizes, every field offset, the ABI version
    fields = {"b200_effector": _lib.Effector, "b200_sixdof_desc": _lib.Desc, "b200_timings": _lib.Timings}
    src = ['#include <stdio.h>', '#include <stddef.h>', '#include "b200_sixdof.h"', 'int main(void) {',
           'printf("abi %u\\n", B200_SIXDOF_ABI_VERSION);']
    for cname, st in fields.items():
        src.append(f'printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in st._fields_:
            src.append(f'printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    src += ["return 0; }"]
    c = tmp_path / "layout.c"
    c.write_text("\n".join(src))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", str(c), "-I", os.path.join(ROOT, "include"), "-o", str(exe)], check=True)
    got = dict(line.split() for line in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    assert int(got["abi"]) == _lib.ABI_VERSION
    for cname, st in fields.items():
        assert int(got[cname]) == ctypes.sizeof(st), cname
        for fname, _ in st._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(st, fname).offset, (cname, fname)


def test_no_gpu_means_loud_failure_not_cpu_fallback(built_lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert el.device_count() == 0
    with pytest.raises(el.B200Error) as ei:
        el.B200Exec(3, 1, 0.01)
    assert ei.value.code == _lib.ERR_NO_DEVICE
    w = el.World()
    w.spawn(el.Body(), name="e1")
    with pytest.raises(el.B200Error):
        w.build(el.six_dof())


def test_product_never_imports_the_oracle():
    """The product package must not import, link, dlopen or call anything under oracle/
    (comments may cite it as the thing the EXACT mode is bit-identical to)."""
    pkg = os.path.join(ROOT, "elodin_b200")
    banned = [r"^\s*(from|import)\s+oracle", r"libsixdof_oracle", r"\borc_[a-z0-9_]+\s*\(", r"#include\s+[\"<].*oracle",
              r"sixdof_oracle\.h", r"importlib.*oracle", r"[\"']oracle[\"'/]"]
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")) or f == "Makefile":
                src = open(os.path.join(dirpath, f), errors="replace").read()
                for pat in banned:
                    assert not re.search(pat, src, flags=re.M), f"{f} matches {pat}"
    # and the shared library does not depend on it
    out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out


def test_header_is_plain_c_and_usable_from_c(built_lib, tmp_path):
    """include/b200_sixdof.h compiles as C99 and a C program links against the library (the
    cgo / FFI view of the boundary).  Without a GPU the program checks the loud failure."""
    exe = tmp_path / "abi_smoke"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", os.path.join(ROOT, "tests", "c", "abi_smoke.c"),
                    "-I", os.path.join(ROOT, "include"), "-L", os.path.join(ROOT, "elodin_b200"), "-lb200_sixdof",
                    "-Wl,-rpath," + os.path.join(ROOT, "elodin_b200"), "-lm", "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "ok" in out.stdout or "failed loudly" in out.stdout


def test_missing_extension_fails_loudly(monkeypatch, tmp_path):
    """No CUDA extension => the product raises; it never substitutes a CPU path."""
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "libb200_sixdof.so"))
    with pytest.raises(el.B200Error, match="no CPU fallback"):
        el.B200Exec(1, 1, 0.01)
    w = el.World()
    w.spawn(el.Body(), name="e1")
    with pytest.raises(el.B200Error):
        w.build(el.six_dof())


def test_cpp_host_mirror_compiles_and_runs(built_lib, tmp_path):
    """include/b200_world.hpp (World / Exec / WorldExec, the compiled-language mirror of the Rust
    executor seam) builds with g++ -std=c++17 and runs the reference's checks through the C ABI; without
    a GPU the program verifies the loud NO_DEVICE failure."""
    exe = tmp_path / "world_exec_test"
    subprocess.run(["g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", os.path.join(ROOT, "tests", "cpp", "world_exec_test.cpp"),
                    "-I", os.path.join(ROOT, "include"), "-L", os.path.join(ROOT, "elodin_b200"), "-lb200_sixdof",
                    "-Wl,-rpath," + os.path.join(ROOT, "elodin_b200"), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "ok" in out.stdout or "failed loudly" in out.stdout
