"""Device-bound inputs at full size: python tools/gpu_devpath.py [workload]  (development probe)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import __graft_entry__ as ge
import bench
pkg = ge.load_package()
name = sys.argv[1] if len(sys.argv) > 1 else "merger64m"
p, e0, mh, desc = bench.make_particles(pkg, name)
n = len(p["x"])
dev = torch.device("cuda", 0)
ctx = pkg.Context(0, 8)
f8 = ["x", "y", "z", "vx", "vy", "vz", "mass", "U", "next_time", "mu"]
t = {k: torch.from_numpy(np.ascontiguousarray(p[k])).to(dev) for k in f8}
t["type"] = torch.from_numpy(np.ascontiguousarray(p["type"])).to(dev)
torch.cuda.synchronize()
ptrs = {k: v.data_ptr() for k, v in t.items()}
for it in range(3):
    try:
        ctx.set_particles_device(ptrs, n)
        R = ctx.build_tree()
        print("separate build", it, R, ctx.counters()["n_nodes"], flush=True)
        ctx.visual_density(R / 1e5); ctx.gas_density(mh); ctx.forces(0.0, e0, 0.5)
        print("  forces ok", ctx.phase_ms(), flush=True)
    except Exception as e:  # noqa: BLE001
        print("separate FAILED", it, e, flush=True)
for it in range(3):
    try:
        ctx.set_particles_device(ptrs, n)
        R = ctx.force_path(R / 1e5, mh, 0.0, e0, 0.5)
        print("force_path", it, R, ctx.phase_ms(), flush=True)
    except Exception as e:  # noqa: BLE001
        print("force_path FAILED", it, e, flush=True)
