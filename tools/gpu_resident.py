"""Device-resident loop probe: python tools/gpu_resident.py workload separate|fused [steps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import __graft_entry__ as ge
import bench
pkg = ge.load_package()
name, mode = sys.argv[1], sys.argv[2]
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 6
p, e0, mh, desc = bench.make_particles(pkg, name)
ctx = pkg.Context(0, 8)
ctx.set_particles(p)
ctx.integrator_init(2.0, 1e13, 1e13, 70.0, e0)
R = ctx.build_tree(); ctx.visual_density(R / 100000); ctx.gas_density(mh); ctx.forces(0.0, e0, 0.5)
ctx.integrator_assign_all()
print("init R", R, flush=True)
for it in range(steps):
    try:
        t = ctx.step_begin()
        if mode == "separate":
            R2 = ctx.build_tree(); ctx.visual_density(R / 100000); ctx.gas_density(mh); ctx.forces(t, e0, 0.5)
        else:
            R2 = ctx.force_path(R / 100000, mh, t, e0, 0.5)
        ctx.step_end()
        c = ctx.counters()
        print(it, "t", t, "R", R2, "outliers", c["n_outliers"], "nodes", c["n_nodes"], "depth", c["max_depth"], flush=True)
    except Exception as e:  # noqa: BLE001
        print(it, "FAILED", e, flush=True)
        st = ctx.state()
        for k in ("x", "vx"):
            a = st[k]; print(k, "nan", int(np.isnan(a).sum()), "inf", int(np.isinf(a).sum()), "absmax", float(np.nanmax(np.abs(a))))
        break
