"""Walk time when only a fraction of the particles is active (individual time steps)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import __graft_entry__ as ge
import bench
pkg = ge.load_package()
p, e0, mh, desc = bench.make_particles(pkg, "plummer1m")
ctx = pkg.Context(0, 8)
rng = np.random.default_rng(1)
for frac in (1.0, 0.5, 0.1, 0.01):
    q = dict(p)
    q["next_time"] = np.where(rng.uniform(0, 1, len(p["x"])) < frac, 0.0, 5.0)
    best = 1e9
    for _ in range(3):
        ctx.set_particles(q); R = ctx.build_tree(); ctx.visual_density(R / 1e5); ctx.gas_density(mh); ctx.forces(0.0, e0, 0.5)
        best = min(best, ctx.phase_ms()["forces"])
    c = ctx.counters()
    print("active fraction %.2f: n_active %d forces %.3f ms interactions %.3e" % (frac, c["n_active"], best, c["interactions"]))
