"""Small mixed workload for compute-sanitizer runs (memcheck / racecheck / initcheck)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
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
    ctx.set_particles(dict(p))                                  # the fused call (one synchronisation) on the same particles
    ctx.force_path(out["R"] / 1e5, mh, 0.0, 1e18, 0.5)
print("done", ctx.counters()["interactions"])
