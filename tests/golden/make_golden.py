"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/ag_ref, built by
oracle/Makefile from /root/reference) — run in the build container only:

    python tests/golden/make_golden.py

Each file holds the inputs (in_*), the parameters (par_*) and the reference's outputs (out_*,
node_*) for one case.  The example ICs are the reference's own shipped files
(input_data/Example/*.dat) read through the reference's own loaders (DataManager::loadICs)."""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import agio, oracle  # noqa: E402
import __graft_entry__ as ge  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def convert(fmt, rel):
    with tempfile.TemporaryDirectory() as d:
        out = os.path.join(d, "p.agp")
        subprocess.check_call([oracle.REF_BIN, "convert", fmt, rel, out], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        return agio.read_agp(out)


def save(name, p, theta, e0, massInH, cores, nodes):
    o = oracle.run_ref(p, theta, e0, massInH, 0.0, cores, nodes=nodes)
    d = {"in_" + k: v for k, v in p.items()}
    d.update({"par_theta": theta, "par_e0": e0, "par_massInH": massInH, "par_cores": cores, "par_globalTime": 0.0})
    for k, v in o.items():
        if k == "nodes":
            if nodes:
                d.update({"node_" + kk: vv for kk, vv in v.items()})
        else:
            d["out_" + k] = v
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
    print(name, len(p["x"]), "R", o["R"])


if __name__ == "__main__":
    assert oracle.have_ref(), "build oracle/_ref first (make -C oracle ref)"
    ics = ge.load_package().ics
    p = convert("gadget", "Example/gassphere_littleendian.dat")
    save("gassphere", p, 0.5, 1e18, 32 * float(p["mass"][0]), 8, True)
    p = convert("makeGal", "Example/galaxy_gas.dat")
    order = np.lexsort((p["z"], p["y"], p["x"]))          # the makeGal loader shuffles with random_device: fix an order
    p = {k: v[order] for k, v in p.items()}
    save("galaxy_gas", p, 0.5, 1e19, 2e39, 8, True)
    p = ics.plummer(3000, seed=42, gas_fraction=0.3)
    save("plummer_gas3k", p, 0.5, 1e18, ics.gas_mass_in_h(p, 16), 2, True)
    p = convert("gadget", "Example/galiC_M1_22k.dat")       # the shipped example of Config.ini:42 (C0)
    save("galic22k", p, 0.5, 1e19, 1e40, 8, False)
    p = convert("gadget", "Example/galaxy_littleendian.dat")  # 60 000 particles (halo + disk), masses from the header
    save("galaxy60k", p, 0.5, 1e19, 1e40, 8, False)
