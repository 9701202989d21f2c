import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import __graft_entry__ as ge
from oracle import oracle
from parity import relerr_vec
pkg = ge.load_package()
ctx = pkg.Context(0, 8)
ctx.set_option(pkg.capi.AGB_OPT_TARGET_COUNTERS, 1)
for variant in ("inactive", "massless", "massless_nogas"):
    p = pkg.ics.plummer(20000, seed=8, gas_fraction=0.0 if variant == "massless_nogas" else 0.2)
    if variant == "inactive":
        p["next_time"][::3] = 7.0
    else:
        p["mass"][5::1000] = 0.0
    mh = pkg.ics.gas_mass_in_h(p, 32)
    got, _ = pkg.run_step(dict(p), 0.5, 1e18, mh, 0.0, context=ctx)
    want = oracle.run(p, 0.5, 1e18, mh, 0.0, 8)
    act = p["next_time"] == 0
    rel = relerr_vec((got["ax"], got["ay"], got["az"]), (want["ax"], want["ay"], want["az"]))
    bad = np.where(act & (rel > 1e-6))[0]
    tc = ctx.target_counters()
    print(variant, "bad", len(bad), "of", act.sum())
    for k in ("visits", "acc_nodes", "acc_leaves", "sph"):
        print("  counter mismatch", k, int((tc[k] != want[k]).sum()))
    print("  h mismatch", int((got["h"] != want["h"]).sum()), "rho maxrel", float(np.max(np.abs(got["rho"] - want["rho"]) / np.where(want["rho"] != 0, want["rho"], 1))))
    for i in bad[:8]:
        print("   i", i, "type", p["type"][i], "m", p["mass"][i], "rel", rel[i], "h", got["h"][i], want["h"][i], "sph", tc["sph"][i], want["sph"][i],
              "accn", tc["acc_nodes"][i], want["acc_nodes"][i], "dUdt", got["dUdt"][i], want["dUdt"][i])
    print("  bad types", np.bincount(p["type"][bad], minlength=4))
