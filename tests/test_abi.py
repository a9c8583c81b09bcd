"""CPU checks of the boundary: the shared library builds, loads, exports every symbol include/fourq_b200.h declares,
and refuses to compute without a GPU (no CPU fallback).  No compute calls are made here."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def libpath():
    from fourq_b200 import build
    return build.build()


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "fourq_b200.h")).read()
    return sorted(set(re.findall(r"FQ_API\s+[\w\s\*]+?\b(fq_\w+)\s*\(", src)))


def test_header_declares_the_boundary():
    syms = declared_symbols()
    for must in ("fq_dh", "fq_dh_affine", "fq_dh_base", "fq_mul_base", "fq_decode", "fq_encode", "fq_fp2_mul", "fq_fp2_sqr",
                 "fq_fp2_inv", "fq_x25519", "fq_dev_run", "fq_imad_peak"):
        assert must in syms


def test_library_exports_every_declared_symbol(libpath):
    L = ctypes.CDLL(libpath)
    for s in declared_symbols():
        assert hasattr(L, s), s
    from fourq_b200 import _lib
    assert sorted(_lib.EXPORTS) == declared_symbols()
    assert _lib.lib().fq_version() == 200


def test_no_cpu_fallback(libpath):
    import fourq_b200
    from fourq_b200 import _lib
    if _lib.lib().fq_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(fourq_b200.FourQError, match="no CPU path"):
        fourq_b200.MUL_base(np.zeros((4, 32), np.uint8))
    with pytest.raises(fourq_b200.FourQError):
        fourq_b200.GFp2.mul(np.zeros((4, 32), np.uint8), np.zeros((4, 32), np.uint8))


def test_argument_validation(libpath):
    import fourq_b200
    with pytest.raises(ValueError):
        fourq_b200.decode(np.zeros((4, 31), np.uint8))          # curve4q.py:50-51 length check
    with pytest.raises(TypeError):
        fourq_b200.decode(np.zeros((4, 32), np.int32))
    with pytest.raises(ValueError):
        fourq_b200.DH(np.zeros((4, 32), np.uint8), np.zeros((5, 32), np.uint8))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "fourq_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                # comments may mention the test oracle; code must not import, include, load or call it
                assert not re.search(r"(import|from|include|CDLL|dlopen)[^\n]*oracle", src), f
                assert not re.search(r"(import|from|CDLL|dlopen)[^\n]*hostsim", src), f


def test_tools_do_not_use_the_oracle():
    """Only tests/ (incl. tests/checks), smoke() and bench.py's checking legs may touch oracle/: the measurement tools must not."""
    tools = os.path.join(ROOT, "tools")
    for dirpath, _, files in os.walk(tools):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle|from oracle|import oracle|c_oracle|fourq_oracle", src, re.M), f


def test_c_example_compiles_links_and_fails_loudly_without_a_gpu(libpath, tmp_path):
    """examples/dh_example.c binds the C ABI from plain C; without a CUDA device it must exit with the no-device error."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    exe = str(tmp_path / "dh_example")
    libdir = os.path.dirname(libpath)
    subprocess.check_call(["gcc", "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-o", exe,
                           os.path.join(ROOT, "examples", "dh_example.c"), "-L", libdir, "-lfourq_b200", "-Wl,-rpath," + libdir])
    from fourq_b200 import _lib
    r = subprocess.run([exe], capture_output=True, text=True)
    if _lib.lib().fq_device_count() > 0:
        assert r.returncode == 0 and "shared secrets agree" in r.stdout, r.stderr
    else:
        assert r.returncode == 2 and "no CPU path" in r.stderr
