"""oracle/agio.py — TEST INFRASTRUCTURE ONLY.

numpy readers/writers for the two little-endian files exchanged with oracle/_ref/ag_ref
(layout documented at the top of oracle/ref_harness.cpp)."""
import numpy as np

PART_FIELDS = ("x", "y", "z", "vx", "vy", "vz", "mass", "U", "next_time", "rho", "P", "T", "mu")


def write_agp(path, p):
    n = len(p["x"])
    with open(path, "wb") as f:
        f.write(b"AGPART01")
        f.write(np.int64(n).tobytes())
        for k in PART_FIELDS:
            f.write(np.ascontiguousarray(p[k], dtype="<f8").tobytes())
        f.write(np.ascontiguousarray(p["type"], dtype=np.uint8).tobytes())


def read_agp(path):
    with open(path, "rb") as f:
        assert f.read(8) == b"AGPART01"
        n = int(np.frombuffer(f.read(8), dtype="<i8")[0])
        p = {k: np.frombuffer(f.read(8 * n), dtype="<f8").copy() for k in PART_FIELDS}
        p["type"] = np.frombuffer(f.read(n), dtype=np.uint8).copy()
    return p


def read_ago(path):
    with open(path, "rb") as f:
        assert f.read(8) == b"AGOUT001"
        n = int(np.frombuffer(f.read(8), dtype="<i8")[0])
        o = {"R": float(np.frombuffer(f.read(8), dtype="<f8")[0])}
        for k in ("ax", "ay", "az", "dUdt", "h", "rho", "P", "T", "vis"):
            o[k] = np.frombuffer(f.read(8 * n), dtype="<f8").copy()
        o["leafdepth"] = np.frombuffer(f.read(4 * n), dtype="<i4").copy()
        o["key_hi"] = np.frombuffer(f.read(8 * n), dtype="<u8").copy()
        o["key_lo"] = np.frombuffer(f.read(8 * n), dtype="<u8").copy()
        for k in ("visits", "acc_nodes", "acc_leaves", "sph"):
            o[k] = np.frombuffer(f.read(4 * n), dtype="<i4").copy()
        m = int(np.frombuffer(f.read(8), dtype="<i8")[0])
        nd = {}
        nd["depth"] = np.frombuffer(f.read(4 * m), dtype="<i4").copy()
        nd["isLeaf"] = np.frombuffer(f.read(4 * m), dtype="<i4").copy()
        nd["nchild"] = np.frombuffer(f.read(8 * m), dtype="<i8").copy()
        nd["key_hi"] = np.frombuffer(f.read(8 * m), dtype="<u8").copy()
        nd["key_lo"] = np.frombuffer(f.read(8 * m), dtype="<u8").copy()
        for k in ("mass", "comx", "comy", "comz", "gasMass", "mvx", "mvy", "mvz"):
            nd[k] = np.frombuffer(f.read(8 * m), dtype="<f8").copy()
        o["nodes"] = nd
    return o
