"""Quick perf + stats run: python tools/gpu_quick.py [workload] [reps]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import __graft_entry__ as ge
import bench
pkg = ge.load_package()
name = sys.argv[1] if len(sys.argv) > 1 else "plummer1m"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
p, e0, mh, desc = bench.make_particles(pkg, name)
if os.environ.get("NOGAS"):
    p["type"][p["type"] == 2] = 1
ctx = pkg.Context(0, 8)
ctx.set_particles(p)
best = None
for i in range(reps):
    ctx.set_particles(p)
    R = ctx.build_tree(); ctx.visual_density(R / 1e5); ctx.gas_density(mh); ctx.forces(0.0, e0, 0.5)
    ph = ctx.phase_ms()
    if best is None or ph["walk_kernel"] < best["walk_kernel"]:
        best = ph
c = ctx.counters()
print(name, json.dumps(best))
print(json.dumps(c))
n = c["n_particles"]
print("interactions/target %.1f  popped/target %.1f  straddling/popped %.3f  popped/round %.2f  straddling/round %.2f opened/round %.2f tiles/group %.1f  Ginter/s %.1f" % (
    c["interactions"] / n, c["walk_popped"] / n, c["walk_straddling"] / max(1, c["walk_popped"]), c["walk_popped"] / max(1, c["walk_rounds"]),
    c["walk_straddling"] / max(1, c["walk_rounds"]), c["walk_opened"] / max(1, c["walk_rounds"]), c["walk_tiles"] / c["groups"], c["interactions"] / best["walk_kernel"] / 1e6))
