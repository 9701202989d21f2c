import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def pkg():
    import __graft_entry__ as ge
    return ge.load_package()


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as o
    o.build()
    return o


GOLDEN = os.path.join(ROOT, "tests", "golden")


def load_golden(name):
    import numpy as np
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    p = {k[3:]: z[k] for k in z.files if k.startswith("in_")}
    out = {k[4:]: z[k] for k in z.files if k.startswith("out_")}
    nodes = {k[5:]: z[k] for k in z.files if k.startswith("node_")}
    if nodes:
        out["nodes"] = nodes
    par = {k[4:]: z[k].item() for k in z.files if k.startswith("par_")}
    return p, out, par


GOLDEN_CASES = ["gassphere", "galaxy_gas", "plummer_gas3k", "galic22k", "galaxy60k"]
