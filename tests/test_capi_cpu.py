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


def test_ctypes_mirror_matches_the_header(pkg, tmp_path):
    """The ctypes structures of capi.py must have the layout a C compiler gives the structs of include/agb200.h
    (a field added on one side only would silently shift every later field)."""
    import subprocess
    fields = {"agb_particles": pkg.capi.Particles, "agb_results": pkg.capi.Results, "agb_aos_layout": pkg.capi.AosLayout, "agb_counters": pkg.capi.Counters}
    prog = ['#include <stdio.h>', '#include <stddef.h>', '#include "agb200.h"', "int main(void) {"]
    for cname, ct in fields.items():
        prog.append('printf("%s %%zu\\n", sizeof(%s));' % (cname, cname))
        for fname, _ in ct._fields_:
            prog.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (cname, fname, cname, fname))
    prog += ["return 0; }"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(prog))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = dict(line.split() for line in subprocess.check_output([str(exe)], text=True).splitlines())
    for cname, ct in fields.items():
        assert int(got[cname]) == C.sizeof(ct), cname
        for fname, _ in ct._fields_:
            assert int(got["%s.%s" % (cname, fname)]) == getattr(ct, fname).offset, (cname, fname)


def test_constants_match_the_header(pkg, tmp_path):
    """Every AGB_* integer constant of capi.py has the value the header's enums give it (a C compiler evaluates them), and every
    AGB_OPT_* / AGB_MEM_* the header declares is mirrored."""
    import subprocess
    hdr = open(os.path.join(ROOT, "include", "agb200.h")).read()
    declared = sorted(set(re.findall(r"\b(AGB_(?:OPT|MEM)_[A-Z_0-9]+)\b\s*=", hdr)))
    mirrored = sorted(k for k in dir(pkg.capi) if re.fullmatch(r"AGB_(OPT|MEM)_[A-Z_0-9]+", k))
    assert declared == mirrored
    prog = ['#include <stdio.h>', '#include "agb200.h"', "int main(void) {"]
    prog += ['printf("%s %%d\\n", (int)%s);' % (k, k) for k in declared]
    prog += ["return 0; }"]
    src = tmp_path / "consts.c"
    src.write_text("\n".join(prog))
    exe = tmp_path / "consts"
    subprocess.check_call(["gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = dict(line.split() for line in subprocess.check_output([str(exe)], text=True).splitlines())
    for k in declared:
        assert int(got[k]) == getattr(pkg.capi, k), k
