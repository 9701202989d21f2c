"""debug: the array-of-structs hand-over (agb_set_particles_aos / agb_get_results_aos) against the SoA hand-over on a golden set,
in-process (synthetic 160-byte records) and through oracle/_ref/ag_ref_gpu (the reference's Particle)"""
import ctypes as C
import os, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import __graft_entry__ as ge
from conftest import load_golden
from oracle import agio
pkg = ge.load_package()
capi = pkg.capi
name = sys.argv[1] if len(sys.argv) > 1 else "galic22k"
p, want, par = load_golden(name)
n = len(p["x"])


def rel(got, ref):
    a = np.stack([got["ax"], got["ay"], got["az"]]); b = np.stack([ref["ax"], ref["ay"], ref["az"]])
    return np.linalg.norm(a - b, axis=0) / np.maximum(np.linalg.norm(b, axis=0), 1e-300)


def report(tag, got, ref, order=None):
    r = rel(got, ref)
    bad = np.flatnonzero(r > 1e-4)
    print("%s: median %.3e p99 %.3e max %.3e bad %d" % (tag, np.median(r), np.percentile(r, 99), r.max(), len(bad)))
    if len(bad):
        print("   bad caller idx:", bad[:12], "types", np.unique(p["type"][bad], return_counts=True))
        if order is not None:
            pos = np.empty(n, np.int64); pos[order] = np.arange(n)
            tp = np.sort(pos[bad])
            runs = np.split(tp, np.flatnonzero(np.diff(tp) != 1) + 1)
            print("   tree positions: %d runs; first runs (start, len): %s" % (len(runs), [(int(q[0]), len(q)) for q in runs[:12]]))
            print("   start %% 32: %s   start %% 256: %s" % ([int(q[0]) % 32 for q in runs[:12]], [int(q[0]) % 256 for q in runs[:12]]))
        i = bad[0]
        print("   particle %d: got (%g %g %g) ref (%g %g %g)" % (i, got["ax"][i], got["ay"][i], got["az"][i], ref["ax"][i], ref["ay"][i], ref["az"][i]))
    return bad


ctx = pkg.Context(0, int(par["cores"]))
soa, _ = pkg.run_step(dict(p), par["theta"], par["e0"], par["massInH"], par["globalTime"], context=ctx)
order = ctx.slice_results(0, 1, names=())["index"].astype(np.int64)
report("SoA vs golden", soa, want, order)

# ---- in-process AoS with a synthetic record
F = ("position", "velocity", "acc", "mass", "type", "U", "next_time", "mu", "rho", "P", "T", "h", "dUdt", "visualDensity")
OFF = dict(position=0, velocity=24, acc=48, mass=72, type=80, U=88, next_time=96, mu=104, rho=112, P=120, T=128, h=136, dUdt=144, visualDensity=152)
STRIDE = 160
buf = np.zeros(n * STRIDE, np.uint8)
f8 = buf.view(np.float64).reshape(n, STRIDE // 8)
f8[:, 0] = p["x"]; f8[:, 1] = p["y"]; f8[:, 2] = p["z"]; f8[:, 3] = p["vx"]; f8[:, 4] = p["vy"]; f8[:, 5] = p["vz"]
f8[:, 9] = p["mass"]; f8[:, 11] = p["U"]; f8[:, 12] = p["next_time"]; f8[:, 13] = p["mu"]; f8[:, 14] = p["rho"]; f8[:, 15] = p["P"]; f8[:, 16] = p["T"]
buf.reshape(n, STRIDE)[:, 80] = p["type"]
ptrs = (C.c_void_p * n)(*[buf.ctypes.data + i * STRIDE for i in range(n)])
L = capi.AosLayout(**OFF)
for rep in range(2):
    c2 = pkg.Context(0, int(par["cores"]))
    lib = c2.lib
    capi.check(c2.h, lib.agb_set_particles_aos(c2.h, ptrs, n, C.byref(L)))
    R = C.c_double()
    capi.check(c2.h, lib.agb_build_tree(c2.h, C.byref(R)))
    capi.check(c2.h, lib.agb_visual_density(c2.h, R.value / 100000))
    capi.check(c2.h, lib.agb_gas_density(c2.h, par["massInH"]))
    capi.check(c2.h, lib.agb_forces(c2.h, par["globalTime"], par["e0"], par["theta"]))
    capi.check(c2.h, lib.agb_get_results_aos(c2.h, ptrs, n, C.byref(L)))
    aos = {"ax": f8[:, 6].copy(), "ay": f8[:, 7].copy(), "az": f8[:, 8].copy()}
    report("in-process AoS (rep %d) vs SoA" % rep, aos, soa, order)
    print("   R equal:", R.value == soa["R"], "counters", {k: c2.counters()[k] for k in ("n_nodes", "n_outliers", "interactions")}, "vs", {k: ctx.counters()[k] for k in ("n_nodes", "n_outliers", "interactions")})
    c2.close()

# ---- the reference's driver code on the GPU tree
BIN = os.path.join(ROOT, "oracle", "_ref", "ag_ref_gpu")
with tempfile.TemporaryDirectory() as d:
    agio.write_agp(os.path.join(d, "in.agp"), p)
    for thr in (int(par["cores"]), 1):
        env = dict(os.environ, OMP_NUM_THREADS=str(thr))
        args = [BIN, "run", os.path.join(d, "in.agp"), os.path.join(d, "out.ago"), repr(par["theta"]), repr(par["e0"]), repr(par["massInH"]), repr(par["globalTime"]), str(int(par["cores"])), "0"]
        subprocess.check_call(args, env=env, stdout=subprocess.DEVNULL)
        got = agio.read_ago(os.path.join(d, "out.ago"))
        report("ag_ref_gpu OMP=%d vs SoA" % thr, got, soa, order)
        print("   R equal:", got["R"] == soa["R"], " vis equal:", np.array_equal(got["vis"], soa["vis"]))
