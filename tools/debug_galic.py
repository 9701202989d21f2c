"""debug: galic22k golden through the SoA path (both precisions) and through ag_ref_gpu (AoS path); prints where they differ"""
import os, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import __graft_entry__ as ge
from conftest import load_golden
from oracle import agio
pkg = ge.load_package()
name = sys.argv[1] if len(sys.argv) > 1 else "galic22k"
p, want, par = load_golden(name)
def rel(got):
    a = np.stack([got["ax"], got["ay"], got["az"]]); b = np.stack([want["ax"], want["ay"], want["az"]])
    return np.linalg.norm(a - b, axis=0) / np.maximum(np.linalg.norm(b, axis=0), 1e-300)
for mixed in (1, 0):
    ctx = pkg.Context(0, int(par["cores"]))
    ctx.set_option(pkg.capi.AGB_OPT_TARGET_COUNTERS, 1)
    ctx.set_option(pkg.capi.AGB_OPT_PRECISION, mixed)
    got, _ = pkg.run_step(dict(p), par["theta"], par["e0"], par["massInH"], par["globalTime"], context=ctx)
    r = rel(got)
    tc = ctx.target_counters()
    bad = np.flatnonzero(r > 1e-4)
    print("SoA mixed=%d: median %.2e p99 %.2e max %.2e, bad %d, count mismatches %s, counters %s" % (mixed, np.median(r), np.percentile(r, 99), r.max(), len(bad),
          {k: int((tc[k] != want[k]).sum()) for k in tc}, {k: ctx.counters()[k] for k in ("n_nodes", "max_depth", "n_outliers", "interactions")}))
    if len(bad):
        print("  first bad:", bad[:10], r[bad[:10]], "leafdepth", want["leafdepth"][bad[:10]])
    ctx.close()
BIN = os.path.join(ROOT, "oracle", "_ref", "ag_ref_gpu")
with tempfile.TemporaryDirectory() as d:
    agio.write_agp(os.path.join(d, "in.agp"), p)
    env = dict(os.environ, OMP_NUM_THREADS=str(int(par["cores"])))
    args = [BIN, "run", os.path.join(d, "in.agp"), os.path.join(d, "out.ago"), repr(par["theta"]), repr(par["e0"]), repr(par["massInH"]), repr(par["globalTime"]), str(int(par["cores"])), "0"]
    for rep in range(2):
        subprocess.check_call(args, env=env, stdout=subprocess.DEVNULL)
        got = agio.read_ago(os.path.join(d, "out.ago"))
        r = rel(got)
        bad = np.flatnonzero(r > 1e-4)
        print("AoS run %d: median %.2e p99 %.2e max %.2e, bad %d" % (rep, np.median(r), np.percentile(r, 99), r.max(), len(bad)), bad[:10], r[bad[:10]])
    if len(sys.argv) > 2:
        out = subprocess.run(["compute-sanitizer", "--tool", sys.argv[2]] + args, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True).stdout
        print(out[-3000:])
