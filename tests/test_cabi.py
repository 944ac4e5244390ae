"""The C-ABI library loads in the build container and exports every symbol include/phz.h declares."""
import ctypes
import os
import re

import pytest

from phaser_b200 import engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "phz.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(phz_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    if not os.path.exists(engine.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    lib = ctypes.CDLL(engine.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), n
    assert set(names) == set(engine.EXPORTS)
    lib.phz_backend_name.restype = ctypes.c_char_p
    assert lib.phz_backend_name() == b"cuda-sm_100a"


def test_engine_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(engine.PhzError):
        engine.Engine(device="cuda:0")
    with pytest.raises(engine.PhzError):
        engine.Engine(device="cpu")


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "phaser_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), fn
            assert "hostsim.cpp" not in src
