"""Lane-span statistics of the walk's interaction lists (counter mode): python tools/gpu_masks.py [workload ...]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
import bench
pkg = ge.load_package()
for name in (sys.argv[1:] or ["plummer1m"]):
    p, e0, mh, desc = bench.make_particles(pkg, name)
    ctx = pkg.Context(0, 8)
    ctx.set_option(pkg.capi.AGB_OPT_TARGET_COUNTERS, 1)
    ctx.set_particles(p)
    R = ctx.build_tree(); ctx.visual_density(R / 1e5); ctx.gas_density(mh); ctx.forces(0.0, e0, 0.5)
    c = ctx.counters()
    g = c["groups"]
    ent = [c["walk_ent_wide"], c["walk_ent_half"], c["walk_ent_quarter"]]
    bits = [c["walk_bits_wide"], c["walk_bits_half"], c["walk_bits_quarter"]]
    print(name, "groups", g, "entries/group %.1f" % (sum(ent) / g), "far entries/group %.1f" % (c["walk_ent_far"] / g),
          "interactions/target %.1f" % (c["interactions"] / c["n_particles"]))
    for k, nm in enumerate(("wide", "half", "quarter")):
        print("  %-8s entries/group %7.1f  bits/entry %5.2f  share of pairs %.3f" % (nm, ent[k] / g, bits[k] / max(1, ent[k]), bits[k] / max(1, sum(bits))))
    print("  evaluation classes (entries/group): far+all %.1f  far %.1f  near %.1f" % (c["walk_ent_class0"] / g, c["walk_ent_class1"] / g, c["walk_ent_class2"] / g))
    print(json.dumps({k: v for k, v in c.items() if k.startswith("walk_")}))
    del ctx
