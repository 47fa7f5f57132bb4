"""CPU-side checks of the drop-in boundary: the product library loads, exports every symbol that
include/mv.h declares, refuses to run without a device (no CPU fallback), and the C++ mirror of the
reference class compiles and links against it. No compute call is made here."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "mv.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mv_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(product_lib):
    lib = ctypes.CDLL(product_lib)
    names = _declared()
    assert len(names) >= 45
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_python_binding_table_matches_header(product_lib):
    from multivolumes_b200 import caster
    b = caster.binding()
    from multivolumes_b200._abi import _COMMON
    bound = {b.prefix + k for k in list(caster._EXTRA) + list(_COMMON)}
    assert set(_declared()) <= bound, sorted(set(_declared()) - bound)


def test_no_cpu_fallback_without_device(product_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from multivolumes_b200 import MultiRayCaster
    with pytest.raises(RuntimeError, match="no CUDA device|no CPU path"):
        MultiRayCaster(grid_size=32, num_volumes=1, width=64, height=36)


def test_product_never_references_the_oracle():
    """Nothing under multivolumes_b200/ may include, link, load or import anything from oracle/."""
    pkg = os.path.join(ROOT, "multivolumes_b200")
    banned = ("libmv_oracle", "oracle_binding", "mvo.h", "mvo_core.h", "mvo_math.h", "mvo_sampler.h", "../oracle", "oracle/_build")
    for dirpath, _, files in os.walk(pkg):
        if "_build" in dirpath or "__pycache__" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                hits = [b for b in banned if b in text]
                assert not hits, (f, hits)
                if f.endswith((".cu", ".cuh", ".h")):
                    assert "mvo_" not in text, f


def test_cxx_mirror_compiles_and_links(product_lib, tmp_path):
    exe = str(tmp_path / "cxx_check")
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cxx_surface_check.cpp"),
                           "-o", exe, product_lib, "-Wl,-rpath," + os.path.dirname(product_lib)])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "init=" in out.stdout


def test_python_binding_arity_matches_header(product_lib):
    """Every prototype of include/mv.h against the ctypes table of the Python mirror: same number of parameters (a drifted
    binding would otherwise pass garbage through the C-ABI without any error)."""
    from multivolumes_b200 import caster
    from multivolumes_b200._abi import _COMMON
    src = open(os.path.join(ROOT, "include", "mv.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    protos = dict(re.findall(r"\b(mv_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S))
    table = dict(_COMMON); table.update(caster._EXTRA)
    checked = 0
    for name, (restype, argtypes) in table.items():
        params = protos.get("mv_" + name)
        if params is None:
            continue
        params = params.strip()
        n = 0 if params in ("", "void") else params.count(",") + 1
        assert n == len(argtypes), ("mv_" + name, params, len(argtypes))
        checked += 1
    assert checked >= 50
