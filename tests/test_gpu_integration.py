"""GPU test of the drop-in boundary: the reference's own driver-side code (harness calling the five Tree
methods exactly like Simulation.cpp:121-139) linked against integration/Tree_agb200.cpp instead of the
reference's Tree.cpp, i.e. the reference running on the GPU path through the C ABI with its own
array-of-structs `std::vector<Particle*>`.  The binary is prebuilt by oracle/Makefile (it needs the
reference headers, which exist only in the build container) and travels to the GPU box."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "ag_ref_gpu")


@pytest.mark.parametrize("name", ["gassphere", "plummer_gas3k", "galic22k"])
def test_reference_driver_on_gpu_tree(name):
    if not os.access(BIN, os.X_OK):
        pytest.skip("oracle/_ref/ag_ref_gpu not built (reference headers absent)")
    from oracle import agio
    p, want, par = load_golden(name)
    with tempfile.TemporaryDirectory() as d:
        agio.write_agp(os.path.join(d, "in.agp"), p)
        env = dict(os.environ, OMP_NUM_THREADS=str(int(par["cores"])))
        subprocess.check_call([BIN, "run", os.path.join(d, "in.agp"), os.path.join(d, "out.ago"), repr(par["theta"]), repr(par["e0"]),
                               repr(par["massInH"]), repr(par["globalTime"]), str(int(par["cores"])), "0"], env=env,
                              stdout=subprocess.DEVNULL)
        got = agio.read_ago(os.path.join(d, "out.ago"))
    assert got["R"] == want["R"]
    gas = p["type"] == 2
    assert np.array_equal(got["h"][gas], want["h"][gas])
    for k in ("rho", "P", "T", "vis"):
        assert np.allclose(got[k], want[k], rtol=1e-12, atol=0), k
    a = np.stack([got["ax"], got["ay"], got["az"]]); b = np.stack([want["ax"], want["ay"], want["az"]])
    rel = np.linalg.norm(a - b, axis=0) / np.maximum(np.linalg.norm(b, axis=0), 1e-300)
    assert np.median(rel) <= 1e-6 and np.percentile(rel, 99) <= 1e-4, (np.median(rel), np.percentile(rel, 99))
    nz = want["dUdt"] != 0
    assert np.array_equal(got["dUdt"] != 0, nz)
    if nz.any():
        du = np.abs(got["dUdt"][nz] - want["dUdt"][nz]) / np.abs(want["dUdt"][nz])
        assert np.median(du) <= 1e-6 and np.percentile(du, 99) <= 1e-4, (np.median(du), np.percentile(du, 99))
