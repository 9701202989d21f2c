"""Diagnostic run on the GPU box: parity report + phase timings for a few cases. Writes gpurun_out/probe.json."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import __graft_entry__ as ge
from oracle import oracle
from parity import compare

pkg = ge.load_package()
ics = pkg.ics
out = {}
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
ctx = pkg.Context(0, 8)
ctx.set_option(pkg.capi.AGB_OPT_TARGET_COUNTERS, 1)
cases = [("plummer_gas_20k", ics.plummer(20000, seed=3, gas_fraction=0.2), 0.5, 1e18, 32),
         ("disk_100k", ics.disk_galaxy(100000, seed=5), 0.5, 1e18, 64)]
for name, p, theta, e0, nb in cases:
    mh = ics.gas_mass_in_h(p, nb)
    try:
        got, _ = pkg.run_step(dict(p), theta, e0, mh, 0.0, context=ctx)
        want = oracle.run(p, theta, e0, mh, 0.0, 8)
        rep = compare(got, want, p, ctx)
        rep["phase_ms"] = ctx.phase_ms(); rep["counters"] = ctx.counters()
    except Exception as e:  # noqa: BLE001
        rep = {"error": repr(e)}
    out[name] = rep
    print(name, json.dumps(rep, default=str), flush=True)
ctx.set_option(pkg.capi.AGB_OPT_TARGET_COUNTERS, 0)
for n in (1000000, 4000000):
    p = ics.plummer(n, seed=1234)
    try:
        for rep_i in range(3):
            t0 = time.perf_counter()
            got, _ = pkg.run_step(dict(p), 0.5, 1e18, 1e40, 0.0, context=ctx)
            t1 = time.perf_counter()
        rep = {"wall_s": t1 - t0, "phase_ms": ctx.phase_ms(), "counters": ctx.counters()}
    except Exception as e:  # noqa: BLE001
        rep = {"error": repr(e)}
    out["plummer_%d" % n] = rep
    print(n, json.dumps(rep, default=str), flush=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "probe.json"), "w"), indent=1, default=str)
