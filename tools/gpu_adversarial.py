"""experiment: the adversarial random sets of tests/parity.py (without lattice points) through the GPU path, both precisions"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import __graft_entry__ as ge
from parity import adversarial_set, assert_parity, compare
from oracle import oracle
pkg = ge.load_package()
bad = 0
ctxs = {}
for seed in range(16):
    p, (theta, e0, mh, gt, cores) = adversarial_set(pkg, seed, lattice=False)
    n = len(p["x"])
    if n < 20:
        continue
    want = oracle.run(p, theta, e0, mh, gt, cores)
    for mixed in (True, False):
        key = (cores, mixed)
        if key not in ctxs:
            c = pkg.Context(0, cores)
            c.set_option(pkg.capi.AGB_OPT_TARGET_COUNTERS, 1)
            c.set_option(pkg.capi.AGB_OPT_PRECISION, 1 if mixed else 0)
            ctxs[key] = c
        ctx = ctxs[key]
        try:
            got, _ = pkg.run_step(dict(p), theta, e0, mh, gt, context=ctx)
            rep = compare(got, want, p, ctx)
            assert_parity(rep)
            print("seed", seed, "n", n, "cores", cores, "mixed" if mixed else "fp64", "ok", "depth", int(want["leafdepth"].max()), "acc_median %.1e" % rep["acc_median"], flush=True)
        except AssertionError as e:
            bad += 1
            print("seed", seed, "n", n, "cores", cores, "mixed" if mixed else "fp64", "FAILED", str(e)[:1500], flush=True)
        except Exception as e:  # noqa: BLE001
            bad += 1
            print("seed", seed, "n", n, "cores", cores, "mixed" if mixed else "fp64", "ERROR", repr(e)[:500], flush=True)
print("failures:", bad)
