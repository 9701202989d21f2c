"""Small mixed workload for compute-sanitizer runs (memcheck / racecheck / initcheck): both precisions, the fused call with the
late-upload path, three-word keys (deep pair + inflated root cube), the device-resident loop with the sub-grid hooks, the
multi-device handle (the same device twice) and the extended-accuracy mode."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import __graft_entry__ as ge
pkg = ge.load_package()
ctx = pkg.Context(0, 8)
ctx.set_option(pkg.capi.AGB_OPT_TARGET_COUNTERS, 1)
for mixed in (1, 0):
    ctx.set_option(pkg.capi.AGB_OPT_PRECISION, mixed)
    p = pkg.ics.disk_galaxy(20000, seed=3)
    mh = pkg.ics.gas_mass_in_h(p, 32)
    out, _ = pkg.run_step(dict(p), 0.5, 1e18, mh, 0.0, context=ctx)
    ctx.tree_particles(); ctx.nodes(); ctx.target_counters()
    for _ in range(2):                                          # the fused call: call by call first, then the late-upload path
        ctx.set_particles(dict(p))
        ctx.force_path(out["R"] / 1e5, mh, 0.0, 1e18, 0.5)
print("parity paths done", ctx.counters()["interactions"])
# three-word keys
q = pkg.ics.plummer(3000, seed=21, gas_fraction=0.3)
R = float(np.abs(np.stack([q["x"], q["y"], q["z"]])).max())
q["x"][2000] = q["x"][100] + R * 2.0 ** -44; q["y"][2000] = q["y"][100]; q["z"][2000] = q["z"][100]
far = np.arange(0, 3000, 60)
q["x"][far] *= 1e3
deep = pkg.Context(0, 8)
out, _ = pkg.run_step(dict(q), 0.5, 1e16, pkg.ics.gas_mass_in_h(q, 16), 0.0, context=deep)
print("deep keys done, max depth", deep.counters()["max_depth"])
deep.close()
# device-resident loop with the hooks
ctx.set_option(pkg.capi.AGB_OPT_PRECISION, 1)
ctx.set_option(pkg.capi.AGB_OPT_COOLING, 1); ctx.set_option(pkg.capi.AGB_OPT_STAR_FORMATION, 7)
p = pkg.ics.plummer(5000, seed=8, gas_fraction=0.3)
mh = pkg.ics.gas_mass_in_h(p, 16)
ctx.set_particles(dict(p)); ctx.integrator_init(0.02, 1e10, 1e13, 70.0, 1e18)
R = ctx.build_tree(); ctx.visual_density(R / 1e5); ctx.gas_density(mh); ctx.forces(0.0, 1e18, 0.5); ctx.integrator_assign_all()
for _ in range(3):
    t = ctx.step_begin(); ctx.force_path(R / 1e5, mh, t, 1e18, 0.5); ctx.step_end()
ctx.state(); ctx.subgrid_state()
print("resident loop done")
# several contexts behind one handle, extended mode
m = pkg.MultiContext([0, 0], 8)
m.set_particles(dict(p)); m.force_path(R / 1e5, mh, 0.0, 1e18, 0.5); m.results()
m.set_option(pkg.capi.AGB_OPT_EXTENDED, 1)
m.set_particles(dict(p)); Rm = m.build_tree(); m.visual_density(Rm / 1e5); m.gas_density(mh); m.forces(0.0, 1e18, 0.5); m.results()
m.close()
print("multi handle + extended mode done")
