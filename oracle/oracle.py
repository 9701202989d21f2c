"""oracle/oracle.py — ctypes front-end of oracle/libag_oracle.so (ag_oracle.c).  TEST INFRASTRUCTURE ONLY.

May be imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference legs."""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libag_oracle.so")
REF_BIN = os.path.join(HERE, "_ref", "ag_ref")
REF_OMP_BIN = os.path.join(HERE, "_ref", "ag_ref_omp")

_pd = C.POINTER(C.c_double)


class _IO(C.Structure):
    _fields_ = [("n", C.c_int64)] + [(k, _pd) for k in ("x", "y", "z", "vx", "vy", "vz", "mass", "U", "next_time", "mu")] + \
        [("type", C.POINTER(C.c_uint8))] + [(k, _pd) for k in ("rho", "P", "T", "ax", "ay", "az", "dUdt", "h", "vis")] + \
        [("leafdepth", C.POINTER(C.c_int32)), ("key_hi", C.POINTER(C.c_uint64)), ("key_lo", C.POINTER(C.c_uint64))] + \
        [(k, C.POINTER(C.c_int32)) for k in ("visits", "acc_nodes", "acc_leaves", "sph")]


def build(force=False):
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(os.path.join(HERE, "ag_oracle.c")):
        subprocess.check_call(["make", "-C", HERE, "libag_oracle.so"], stdout=subprocess.DEVNULL)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
        _lib.ag_oracle_create.restype = C.c_void_p
        _lib.ag_oracle_create.argtypes = [C.POINTER(_IO)]
        _lib.ag_oracle_build.restype = C.c_double
        _lib.ag_oracle_build.argtypes = [C.c_void_p, C.c_int]
        _lib.ag_oracle_visual_density.argtypes = [C.c_void_p, C.c_double]
        _lib.ag_oracle_gas_density.argtypes = [C.c_void_p, C.c_double]
        _lib.ag_oracle_forces.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double]
        _lib.ag_oracle_paths.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        _lib.ag_oracle_nodes.restype = C.c_int64
        _lib.ag_oracle_nodes.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int64),
                                         C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)] + [_pd] * 8
        _lib.ag_oracle_destroy.argtypes = [C.c_void_p]
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def default_particles(n):
    """Particle dict with the reference's defaults (Particle.h:18-57): type 1, mu 0.58, everything else 0."""
    p = {k: np.zeros(n) for k in ("x", "y", "z", "vx", "vy", "vz", "mass", "U", "next_time", "rho", "P", "T")}
    p["mu"] = np.full(n, 0.58)
    p["type"] = np.ones(n, dtype=np.uint8)
    return p


def run(p, theta, e0, massInH, globalTime=0.0, cores=8, visual_radius=None, nodes=True, counters=True, phases=None):
    """Run build -> visual density -> gas density -> forces (Simulation.cpp:120-139) on the C restatement.
    Returns a dict shaped like agio.read_ago()."""
    L = lib()
    n = len(p["x"])
    f8 = lambda k: np.ascontiguousarray(p[k], dtype=np.float64)
    inp = {k: f8(k) for k in ("x", "y", "z", "vx", "vy", "vz", "mass", "U", "next_time", "mu")}
    typ = np.ascontiguousarray(p["type"], dtype=np.uint8)
    o = {k: f8(k).copy() for k in ("rho", "P", "T")}
    for k in ("ax", "ay", "az", "dUdt", "h", "vis"):
        o[k] = f8(k).copy() if k in p else np.zeros(n)       # carried state: acc of inactive particles, accumulated dUdt, h
    o["leafdepth"] = np.zeros(n, dtype=np.int32)
    o["key_hi"] = np.zeros(n, dtype=np.uint64)
    o["key_lo"] = np.zeros(n, dtype=np.uint64)
    for k in ("visits", "acc_nodes", "acc_leaves", "sph"):
        o[k] = np.zeros(n, dtype=np.int32)
    io = _IO()
    io.n = n
    for k in ("x", "y", "z", "vx", "vy", "vz", "mass", "U", "next_time", "mu"):
        setattr(io, k, _p(inp[k], C.c_double))
    io.type = _p(typ, C.c_uint8)
    for k in ("rho", "P", "T", "ax", "ay", "az", "dUdt", "h", "vis"):
        setattr(io, k, _p(o[k], C.c_double))
    if counters:
        for k in ("visits", "acc_nodes", "acc_leaves", "sph"):
            setattr(io, k, _p(o[k], C.c_int32))
    import time
    h = L.ag_oracle_create(C.byref(io))
    try:
        t0 = time.perf_counter()
        o["R"] = L.ag_oracle_build(h, int(cores))
        t1 = time.perf_counter()
        L.ag_oracle_visual_density(h, o["R"] / 100000 if visual_radius is None else visual_radius)
        t2 = time.perf_counter()
        L.ag_oracle_gas_density(h, float(massInH))
        t3 = time.perf_counter()
        L.ag_oracle_forces(h, float(globalTime), float(e0), float(theta))
        t4 = time.perf_counter()
        if phases is not None:
            phases.update(build=t1 - t0, visual=t2 - t1, gas_density=t3 - t2, forces=t4 - t3)
        L.ag_oracle_paths(h, _p(o["leafdepth"], C.c_int32), _p(o["key_hi"], C.c_uint64), _p(o["key_lo"], C.c_uint64))
        if nodes:
            nul = lambda t: C.cast(None, C.POINTER(t))
            m = L.ag_oracle_nodes(h, nul(C.c_int32), nul(C.c_int32), nul(C.c_int64), nul(C.c_uint64), nul(C.c_uint64), *([nul(C.c_double)] * 8))
            nd = {"depth": np.zeros(m, np.int32), "isLeaf": np.zeros(m, np.int32), "nchild": np.zeros(m, np.int64),
                  "key_hi": np.zeros(m, np.uint64), "key_lo": np.zeros(m, np.uint64)}
            for k in ("mass", "comx", "comy", "comz", "gasMass", "mvx", "mvy", "mvz"):
                nd[k] = np.zeros(m)
            L.ag_oracle_nodes(h, _p(nd["depth"], C.c_int32), _p(nd["isLeaf"], C.c_int32), _p(nd["nchild"], C.c_int64),
                              _p(nd["key_hi"], C.c_uint64), _p(nd["key_lo"], C.c_uint64),
                              *[_p(nd[k], C.c_double) for k in ("mass", "comx", "comy", "comz", "gasMass", "mvx", "mvy", "mvz")])
            o["nodes"] = nd
    finally:
        L.ag_oracle_destroy(h)
    return o


def have_ref():
    return os.access(REF_BIN, os.X_OK)


def run_ref_steps(p, theta, e0, massInH, cores, eta, min_ts, max_ts, H0, nsteps):
    """initial forces + nsteps iterations of the reference's main loop on oracle/_ref/ag_ref (its own integrator + Tree)."""
    from . import agio
    with tempfile.TemporaryDirectory() as d:
        agio.write_agp(os.path.join(d, "in.agp"), p)
        out = os.path.join(d, "out.agp")
        subprocess.check_call([REF_BIN, "steps", os.path.join(d, "in.agp"), out] + [repr(float(v)) for v in (theta, e0, massInH)] + [str(int(cores))] +
                              [repr(float(v)) for v in (eta, min_ts, max_ts, H0)] + [str(int(nsteps))], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        q = agio.read_agp(out)
        n = len(q["x"])
        raw = np.fromfile(out + ".acc", dtype="<f8")
        for i, k in enumerate(("ax", "ay", "az", "dUdt", "h", "vis", "next_time2", "timeStep")):
            q[k] = raw[i * n:(i + 1) * n].copy()
        q["globalTime"] = float(raw[8 * n])
    return q


def run_ref(p, theta, e0, massInH, globalTime=0.0, cores=8, nodes=True):
    """Same run on the unmodified reference binary (oracle/_ref/ag_ref); only where it has been built."""
    from . import agio
    with tempfile.TemporaryDirectory() as d:
        agio.write_agp(os.path.join(d, "in.agp"), p)
        subprocess.check_call([REF_BIN, "run", os.path.join(d, "in.agp"), os.path.join(d, "out.ago"), repr(float(theta)), repr(float(e0)),
                               repr(float(massInH)), repr(float(globalTime)), str(int(cores)), "1" if nodes else "0"],
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        return agio.read_ago(os.path.join(d, "out.ago"))
