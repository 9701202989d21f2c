"""SPH pair-kernel statistics of a workload: python tools/gpu_sphstats.py [workload]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import __graft_entry__ as ge
import bench
pkg = ge.load_package()
name = sys.argv[1] if len(sys.argv) > 1 else "gas16m"
p, e0, mh, desc = bench.make_particles(pkg, name)
ctx = pkg.Context(0, 8)
ctx.set_particles(p)
R = ctx.build_tree(); ctx.visual_density(R / 1e5); ctx.gas_density(mh); ctx.forces(0.0, e0, 0.5)
ctx.forces(0.0, e0, 0.5)
c = ctx.counters()
ngas = int((p["type"] == 2).sum())
print(name, "gas", ngas, "orphans", c["gas_orphans"], "records", c["sph_records"], "pairs", c["sph_interactions"],
      "pairs/record %.1f" % (c["sph_interactions"] / max(1, c["sph_records"])), "pairs/gas target %.1f" % (c["sph_interactions"] / max(1, ngas - c["gas_orphans"])),
      "records/group(all) %.2f" % (c["sph_records"] / c["groups"]), "kernel ms", ctx.kernel_ms())
