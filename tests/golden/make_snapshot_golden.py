"""Generates tests/golden/snapshots/: one synthetic `.age` input (written here with numpy) and the `.ag`, `.agc`, `.age`
and `.gadget` files the UNMODIFIED reference writes from it (DataManager::loadICs + DataManager::saveData through
oracle/_ref/ag_ref `save`).  Run in the build container only (needs /root/reference):

    python tests/golden/make_snapshot_golden.py
"""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
OUT = os.path.join(ROOT, "tests", "golden", "snapshots")
REF = os.path.join(ROOT, "oracle", "_ref", "ag_ref")

AGE_REC = np.dtype([("pos", "<f8", 3), ("vel", "<f8", 3), ("mass", "<f8"), ("T", "<f8"), ("P", "<f8"), ("vis", "<f8"), ("U", "<f8"),
                    ("type", "u1"), ("part", "u1"), ("id", "<u4")])
assert AGE_REC.itemsize == 94

N, COUNT, STEP, DT, END, NOW = 240, 200, 7, 2.5e13, 1e16, 1.75e14


def synth(n=N, seed=77):
    rng = np.random.default_rng(seed)
    r = np.zeros(n, dtype=AGE_REC)
    r["pos"] = rng.normal(0, 3e20, (n, 3)); r["vel"] = rng.normal(0, 2e5, (n, 3))
    r["mass"] = rng.uniform(1e35, 3e36, n); r["T"] = rng.uniform(10, 1e6, n); r["P"] = rng.uniform(1e-15, 1e-11, n)
    r["vis"] = rng.uniform(1e-24, 1e-20, n); r["U"] = rng.uniform(1e8, 1e11, n)
    # every (type, galaxyPart) combination, including the ones the Gadget writer only counts
    r["type"] = rng.integers(1, 4, n); r["part"] = rng.integers(1, 4, n)
    r["id"] = rng.permutation(n).astype(np.uint32) + 1000
    return r


def write_age(path, r):
    with open(path, "wb") as f:
        counts = [int((r["type"] == t).sum()) for t in (1, 2, 3)]
        f.write(np.array(counts + [0], dtype="<i4").tobytes())
        f.write(np.array([DT, END, 0.0], dtype="<f8").tobytes())
        f.write(r.tobytes())


def main():
    os.makedirs(OUT, exist_ok=True)
    src = os.path.join(OUT, "in.age")
    write_age(src, synth())
    rel = os.path.relpath(src, "/root/reference/input_data")
    for fmt in ("ag", "agc", "age", "gadget"):
        subprocess.check_call([REF, "save", "age", rel, fmt, OUT + "/", str(STEP), repr(DT), repr(END), repr(NOW), str(COUNT)], stdout=subprocess.DEVNULL)
        os.replace(os.path.join(OUT, "%d.%s" % (STEP, fmt)), os.path.join(OUT, "ref.%s" % fmt))
    # the reference's readers for its two render formats, as .agp column dumps
    for fmt in ("ag", "agc"):
        rel2 = os.path.relpath(os.path.join(OUT, "ref.%s" % fmt), "/root/reference/input_data")
        subprocess.check_call([REF, "convert", fmt, rel2, os.path.join(OUT, "ref_%s_loaded.agp" % fmt)], stdout=subprocess.DEVNULL)
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    sys.exit(main())
