"""CPU tests of the boundary: the C-ABI library builds/loads, exports every symbol that
include/agb200.h declares, and refuses to run without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "agb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(agb_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_exported(pkg):
    lib = pkg.capi.load()
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), n
    assert set(names) == set(pkg.capi.EXPORTS)


def test_no_cpu_fallback(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = pkg.capi.load()
    h = C.c_void_p()
    st = lib.agb_create(C.byref(h), 0, 8)
    assert st == 1 and not h.value                      # AGB_ERR_NO_DEVICE
    with pytest.raises(pkg.capi.AgbError):
        pkg.Context(0, 8)


def test_product_does_not_import_oracle():
    pk = os.path.join(ROOT, "astrogenesis2.0_b200")
    for dirpath, _, files in os.walk(pk):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"(import|from|include|CDLL|dlopen)[^\n]*oracle", txt), os.path.join(dirpath, f)
                assert "libag_oracle" not in txt and "ag_ref" not in txt, os.path.join(dirpath, f)


def test_ics_shapes(pkg):
    p = pkg.ics.disk_galaxy(1000, seed=1)
    assert len(p["x"]) == 1000 and set(p["type"]) <= {1, 2, 3} and (p["type"] == 2).sum() > 0
    p = pkg.ics.plummer(100, seed=1)
    assert p["mass"].sum() == pytest.approx(1e11 * pkg.ics.MSUN)
