"""CPU: the C-ABI library loads here (no GPU), exports every symbol include/b200nuts.h declares, and the
compute entry points fail loudly without a device (there is no CPU fallback in the product path)."""
import ctypes
import os
import re

import numpy as np
import pytest

from pymc_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "b200nuts.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = header_symbols()
    assert "b200_nuts_run" in names and "b200_logp_dlogp" in names and "b200_leapfrog" in names
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/b200nuts.h but not exported"
    assert {s[0] for s in _lib.SYMBOLS} == set(names), "ctypes table and header disagree"
    assert lib.b200_version() == 300


def test_struct_layouts_match_header():
    # field order/size of the ctypes mirrors (a mismatch would corrupt arguments silently)
    assert ctypes.sizeof(_lib.Pcg64State) == 32 and _lib.PCG64_DTYPE.itemsize == 32
    lib = _lib.load()
    assert ctypes.sizeof(_lib.NutsCfg) == lib.b200_struct_size(1) == 184 and ctypes.sizeof(_lib.ChainStateC) == lib.b200_struct_size(13)
    assert ctypes.sizeof(_lib.Stats) == 12 * ctypes.sizeof(ctypes.c_void_p)
    assert ctypes.sizeof(_lib.ChainSummary) == 5 * ctypes.sizeof(ctypes.c_void_p) == lib.b200_struct_size(3)
    assert _lib.ModelDesc.n_obs.offset == 8 and _lib.ModelDesc.x.offset == 24


def test_no_cpu_fallback_without_device():
    lib = _lib.load()
    if lib.b200_device_count() > 0:
        pytest.skip("a GPU is visible")
    from pymc_b200 import engine, models

    with pytest.raises(_lib.B200Error, match="no CUDA device"):
        engine.CompiledModel(models.eight_schools())
    d = _lib.ModelDesc()
    d.kind, d.n = 0, 4
    h = ctypes.c_void_p()
    assert lib.b200_model_create(ctypes.byref(d), ctypes.byref(h)) != 0
    assert b"no CUDA device" in lib.b200_last_error()


def test_product_path_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "pymc_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), fn
