"""debug: device-resident loop with the sub-grid hooks against the numpy restatement, step by step"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import __graft_entry__ as ge
from oracle import oracle, integrator
oracle.build()
pkg = ge.load_package()
p = pkg.ics.plummer(3000, seed=42, gas_fraction=0.3)
for k in ("x", "y", "z"):
    p[k] = p[k] * 0.2
gas = p["type"] == 2
p["U"][gas] = np.where(np.arange(gas.sum()) % 2 == 0, 1e8, p["U"][gas])
mh = pkg.ics.gas_mass_in_h(p, 16)
ctx = pkg.Context(0, 8)
ctx.set_option(pkg.capi.AGB_OPT_PRECISION, 0)
ctx.set_option(pkg.capi.AGB_OPT_COOLING, 1)
ctx.set_option(pkg.capi.AGB_OPT_STAR_FORMATION, 7)
ctx.set_particles(dict(p))
ctx.integrator_init(2.0, 1e13, 2e14, 70.0, 1e18)
R = ctx.build_tree(); ctx.visual_density(R / 100000); ctx.gas_density(mh); ctx.forces(0.0, 1e18, 0.5)
ctx.integrator_assign_all()
for it in range(8):
    want = integrator.run_steps(p, integrator.oracle_forces(0.5, 1e18, mh, 8), e0=1e18, eta=2.0, min_ts=1e13, max_ts=2e14, H0=70.0, nsteps=it + 1, cooling=True, sf_seed=7)
    try:
        t = ctx.step_begin()
        ctx.force_path(R / 100000, mh, t, 1e18, 0.5)
        ctx.step_end()
    except Exception as e:  # noqa: BLE001
        print(it, "FAILED", e)
        st = ctx.state()
        pos = np.stack([st["x"], st["y"], st["z"]], 1)
        u, cnt = np.unique(pos, axis=0, return_counts=True)
        print("  duplicate positions:", int((cnt > 1).sum()), "nan", int(np.isnan(pos).sum()), "inf", int(np.isinf(pos).sum()))
        if (cnt > 1).any():
            d = u[cnt > 1][0]
            idx = np.flatnonzero((pos == d).all(1))
            print("   first duplicate", d, "particles", idx, "types", p["type"][idx], "want pos", [(want["x"][i], want["y"][i], want["z"][i]) for i in idx[:3]])
        break
    st = ctx.state(); res = ctx.results(); ty, sfr = ctx.subgrid_state()
    c = ctx.counters()
    dx = np.abs(st["x"] - want["x"]).max() / np.abs(want["x"]).max()
    print(it, "t", t, want["globalTime"], "depth", c["max_depth"], "dx", dx, "type mismatches", int((ty != want["type"]).sum()), "stars", int((ty != p["type"]).sum()),
          "U maxrel", float(np.nanmax(np.abs(st["U"] - want["U"]) / np.maximum(np.abs(want["U"]), 1e-300))), "nanU", int(np.isnan(st["U"]).sum()), int(np.isnan(want["U"]).sum()),
          "vmax", float(np.abs(st["vx"]).max()), float(np.abs(want["vx"]).max()))
