"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/spsparse_b200.h declares (no compute calls -- those need a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from spsparse_b200 import build, _lib
    build.build()
    return _lib.load()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "spsparse_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(spb_[a-z_0-9]+)\s*\(", text)))


def test_exports_every_declared_symbol(lib):
    from spsparse_b200 import _lib
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/spsparse_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == names, "ctypes binding and header disagree"


def test_version_and_error_string(lib):
    assert lib.spb_version() >= 100
    assert isinstance(lib.spb_last_error(), bytes)


def test_no_gpu_is_a_loud_error_not_a_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = ctypes.c_void_p()
    rc = lib.spb_ctx_create(0, None, ctypes.byref(h))
    assert rc != 0 and b"no CPU fallback" in lib.spb_last_error()


def test_product_does_not_touch_the_oracle():
    """The oracle is test infrastructure: nothing under spsparse_b200/ or include/ may reference it."""
    for base in ("spsparse_b200", "include"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp")):
                    src = open(os.path.join(dp, f), errors="replace").read()
                    assert "liboracle" not in src and "from oracle" not in src and "import oracle" not in src, f
