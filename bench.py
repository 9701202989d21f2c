#!/usr/bin/env python
"""bench.py — headline benchmark of the force hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

A step = one pass of the path over one synthetic particle set: build_tree -> visual_density ->
gas_density -> forces (the reference's Simulation::run step, Simulation.cpp:276-285), all particles
active.  `value` = particle-updates/s with the particles already resident in HBM; `e2e` = the same
through the reference-facing Tree API with pinned HOST arrays (H2D of the particles and D2H of the
results inside the timed region).  N > 1 (torchrun): every rank owns 1/N of the particles, the
positions are all-gathered over NCCL each step, every GPU builds the same tree and walks its own
slice of the tree-ordered targets (strong scaling).

--impl reference times the reference's own CPU implementation (oracle/_ref/ag_ref_omp, compiled in
place from the unmodified sources; all host threads) on the same workload; falls back to the
single-threaded C oracle port when that binary is absent.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "particle-updates/s"
WORKLOADS = {
    # name: (generator, n, e0, neighbours for massInH, description)
    "plummer1m": ("plummer", 1_000_000, 1e18, 0, "C1 Plummer sphere 1M, gravity-only, theta=0.5"),
    "disk4m": ("disk", 4_000_000, 1e18, 64, "C2 disk galaxy 4M (halo+disk+bulge, 25% of disk gas)"),
    "gas16m": ("gasdisk", 16_000_000, 1e18, 64, "C3 gas-rich disk 16M (50% of disk gas)"),
    "merger64m": ("merger", 64_000_000, 1e18, 64, "C4 merger 64M"),
    "plummer100k": ("plummer", 100_000, 1e18, 0, "small Plummer (debug)"),
    "disk400k": ("disk", 400_000, 1e18, 64, "small disk (debug)"),
}
THETA = 0.5


def make_particles(pkg, name, n_override=None):
    gen, n, e0, nb, desc = WORKLOADS[name]
    n = n_override or n
    ics = pkg.ics
    if gen == "plummer":
        p = ics.plummer(n, seed=1234)
    elif gen == "disk":
        p = ics.disk_galaxy(n, seed=1234, gas_disk_fraction=0.25)
    elif gen == "gasdisk":
        p = ics.disk_galaxy(n, seed=1234, gas_disk_fraction=0.5)
    else:
        p = ics.merger(n, seed=1234)
    mh = ics.gas_mass_in_h(p, nb) if nb else 1e40
    return p, e0, mh, desc


# ------------------------------------------------------------------ clocks sampling
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def wait_first(self, timeout=3.0):
        t0 = time.perf_counter()
        while not self.rows and time.perf_counter() - t0 < timeout:
            time.sleep(0.02)

    def stop(self, t_begin=None, t_end=None):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [r for t, r in self.rows if t_begin is None or (t_begin <= t <= t_end + 0.15)]
        window = "timed region"
        if not rows:
            rows, window = [r for _, r in self.rows], "whole run (timed region shorter than the 100 ms sampling period)"
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm), "window": window}


# ------------------------------------------------------------------ reference / CPU baseline
def cpu_reference_run(p, e0, mh, reps, tmpdir):
    """Times the reference path on host cores. Returns (seconds per step list, kind, cores, phases)."""
    from oracle import agio, oracle
    n = len(p["x"])
    if os.access(oracle.REF_OMP_BIN, os.X_OK):
        path = os.path.join(tmpdir, "bench.agp")
        agio.write_agp(path, p)
        out = subprocess.run([oracle.REF_OMP_BIN, "time", path, repr(THETA), repr(e0), repr(mh), "0", str(reps)],
                             stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, check=True).stdout
        rows = [json.loads(l) for l in out.splitlines() if l.startswith("{")]
        secs = [r["build"] + r["visual"] + r["gas_density"] + r["forces"] for r in rows]
        return secs, "reference", os.cpu_count(), rows
    secs, rows = [], []
    for _ in range(reps):
        ph = {}
        oracle.run(p, THETA, e0, mh, 0.0, cores=1, nodes=False, counters=False, phases=ph)
        secs.append(sum(ph.values())); rows.append(ph)
    return secs, "port", 1, rows


def sample_size_for_cpu(n_full, steps, budget_s=25.0):
    # ~1e-5 s per particle-step on 8 cores for the reference; keep the whole leg near budget_s
    cores = os.cpu_count() or 8
    per = 1.0e-5 * 8.0 / max(1, min(cores, 32))
    n = int(budget_s / max(1, steps) / per)
    return max(20_000, min(n_full, n))


def run_reference_arm(args, pkg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    name = args.workload or "plummer1m"
    n_full = WORKLOADS[name][1]
    total = args.steps + args.warmup
    n = sample_size_for_cpu(n_full, total, budget_s=150.0)
    p, e0, mh, desc = make_particles(pkg, name, n)
    with tempfile.TemporaryDirectory() as d:
        secs, kind, cores, rows = cpu_reference_run(p, e0, mh, total, d)
    timed = secs[args.warmup:]
    t = float(np.mean(timed))
    value = n / t
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "particles/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": name, "description": desc, "n_particles": n_full, "theta": THETA, "e0": e0},
        "cpu_baseline": {"value": value, "unit": "particles/s", "cores": cores, "kind": kind,
                         "sample": "%d of %d particles of %s, %d timed steps, phases build+visual+gas_density+forces" % (n, n_full, name, len(timed))},
        "e2e": {"value": value, "unit": "particles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "phases_s": rows[-1],
    }
    print(json.dumps(line))


# ------------------------------------------------------------------ GPU arm
def run_gpu_arm(args, pkg):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    name = args.workload or "plummer1m"
    p, e0, mh, desc = make_particles(pkg, name)
    n = len(p["x"])
    any_gas = bool((p["type"] == 2).any())
    ctx = pkg.Context(local, 8)
    stream = torch.cuda.ExternalStream(ctx.stream(), device=dev)

    f8 = ["x", "y", "z", "vx", "vy", "vz", "mass", "U", "next_time", "mu"] if any_gas else ["x", "y", "z", "mass"]
    # rank-local shard of the particle arrays (what a distributed driver would own and integrate)
    lo, hi = pkg.shard.shard_bounds(n, rank, world)
    shard = {k: torch.from_numpy(np.ascontiguousarray(p[k][lo:hi])).to(dev) for k in f8}
    shard["type"] = torch.from_numpy(np.ascontiguousarray(p["type"][lo:hi])).to(dev)
    full = {k: (torch.empty(n, dtype=v.dtype, device=dev) if world > 1 else v) for k, v in shard.items()}
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)      # > 126 MB L2

    packed = None
    if world > 1 and n % world == 0:
        # every rank's shard lives in its slot of the full-length arrays; one coalesced group of in-place all-gathers per step
        packed = pkg.shard.InPlaceGather({k: v.dtype for k, v in shard.items()}, n, world, rank, dev)
        for k in shard:
            packed.shard[k].copy_(shard[k])
        shard = packed.shard
        full = packed.out
    elif world > 1:
        packed = pkg.shard.PackedGather(shard, n, world)
        full = packed.out

    def gather():
        if world > 1:
            packed.gather(shard)

    breakdown = os.environ.get("AGB_BENCH_BREAKDOWN") == "1"      # development: per-step device time of the exchange vs the path
    bd_ev = []

    def step():
        if breakdown:
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            e[0].record(torch.cuda.current_stream())
        gather()
        if breakdown:
            e[1].record(torch.cuda.current_stream())
        if world > 1:
            stream.wait_stream(torch.cuda.current_stream())
        ptrs = {k: full[k].data_ptr() for k in full}
        ctx.set_particles_device(ptrs, n)
        # build_tree + visual_density + gas_density + forces, one host synchronisation (agb_force_path); the visual-density
        # radius is fixed at init like in the reference (Simulation.cpp:126)
        ctx.force_path(vis_radius, mh, 0.0, e0, THETA, rank, world)
        if breakdown:
            e[2].record(stream)
            bd_ev.append((e, ctx.phase_ms()))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    gather()
    torch.cuda.synchronize()                                                   # the uploads above ran on torch's stream, the path runs on its own
    if world > 1:                                                              # the exchange really delivered every rank's shard (checked once, untimed)
        for k in full:
            if not torch.equal(full[k], torch.from_numpy(np.ascontiguousarray(p[k])).to(dev)):
                raise SystemExit("all-gather of %s does not reproduce the particle set" % k)
    ctx.set_particles_device({k: full[k].data_ptr() for k in full}, n)
    vis_radius = ctx.build_tree() / 100000                                     # Simulation.cpp:123-126
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        sampler.wait_first()
    for _ in range(args.warmup):
        flush.fill_(1)
        step()
    barrier()
    launches0 = ctx.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    walk_ms, build_ms, inter = [], [], 0
    barrier()
    t_wall0 = time.perf_counter()
    for i in range(args.steps):
        flush.fill_(i & 0xff)                      # L2 flush between timed iterations (outside the timed interval)
        barrier()
        ev[i][0].record(stream if world == 1 else torch.cuda.current_stream())
        step()
        ev[i][1].record(stream)
        ev[i][1].synchronize()
        ph = ctx.phase_ms()
        walk_ms.append(ph["walk_kernel"]); build_ms.append(ph["build"])
        inter = ctx.counters()["interactions"]
    barrier()
    t_wall = time.perf_counter() - t_wall0
    step_ms = [a.elapsed_time(b) for a, b in ev]
    launches = ctx.launch_count() - launches0
    clocks = sampler.stop(t_wall0, t_wall0 + t_wall) if rank == 0 else None
    tot = torch.tensor([sum(step_ms), float(inter), sum(walk_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        mx = tot.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = tot.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        total_ms, walk_total_ms, inter_all = float(mx[0]), float(mx[2]), float(sm[1])
    else:
        total_ms, walk_total_ms, inter_all = float(tot[0]), float(tot[2]), float(inter)
    ms_per_step = total_ms / args.steps
    value = n / (ms_per_step * 1e-3)

    # ---- end-to-end through the Tree API with pinned host arrays (N GPUs: every rank uploads its shard, results of its slice come back)
    e2e = None
    if world == 1:
        host = {}
        for k in f8 + ["type"]:
            t = torch.from_numpy(np.ascontiguousarray(p[k])).pin_memory()
            host[k] = t.numpy()
        sim = pkg.Simulation(host, THETA, e0, mh, 0.0)
        tree = pkg.Tree(sim, ctx)
        outn = ("ax", "ay", "az", "visualDensity") + (("dUdt", "h", "rho", "P", "T") if any_gas else ())
        pinned_out = {k: torch.empty(n, dtype=torch.float64).pin_memory() for k in outn}

        out_np = {k: v.numpy() for k, v in pinned_out.items()}
        ctx.bind_results(out_np)          # density outputs leave while the walk runs; acc / dUdt follow in results_into

        def e2e_step():
            ctx.set_particles(host)
            ctx.force_path(vis_radius, mh, 0.0, e0, THETA)
            ctx.results_into(out_np)
        for _ in range(max(1, args.warmup)):
            e2e_step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        torch.cuda.synchronize()
        te = (time.perf_counter() - t0) / args.steps
        h2d = sum(host[k].nbytes for k in host) + 0
        d2h = sum(v.numel() * 8 for v in pinned_out.values())
        e2e = {"value": n / te, "unit": "particles/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "ms_per_step": te * 1e3}
        ctx.bind_results(None)
    else:
        # multi-GPU e2e: host shard -> device shard (H2D), gather, step, then every rank brings back the compact results of
        # ITS slice of the targets (caller-order index + acc [+ dUdt]; agb_get_slice_results) into pinned host memory
        host = {k: torch.from_numpy(np.ascontiguousarray(p[k][lo:hi])).pin_memory() for k in f8 + ["type"]}
        cols = ("ax", "ay", "az") + (("dUdt",) if any_gas else ())
        cap = n // world + 1024
        out_t = {k: torch.empty(cap, dtype=torch.float64).pin_memory() for k in cols}
        out_t["index"] = torch.empty(cap, dtype=torch.int32).pin_memory()
        out_np = {k: v.numpy() for k, v in out_t.items()}
        out_np["index"] = out_np["index"].view(np.uint32)

        def e2e_step():
            for k in host:
                shard[k].copy_(host[k], non_blocking=True)
            step()
            return ctx.slice_results(rank, world, names=cols, out=out_np)
        r = e2e_step()
        mine = len(r["index"]) * (4 + 8 * len(cols))
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        barrier()
        te = torch.tensor([(time.perf_counter() - t0) / args.steps, 0.0], dtype=torch.float64, device=dev)
        by = torch.tensor([0.0, float(mine)], dtype=torch.float64, device=dev)
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        dist.all_reduce(by, op=dist.ReduceOp.SUM)
        te = float(te[0])
        e2e = {"value": n / te, "unit": "particles/s", "h2d_bytes_per_step": int(sum(v.numel() * v.element_size() for v in host.values()) * world),
               "d2h_bytes_per_step": int(by[1]), "ms_per_step": te * 1e3,
               "what": "per rank: H2D of its particle shard, NCCL all-gather, build, densities, walk of its target slice, D2H of that slice's (index, acc[, dUdt])"}

    # ---- device-resident simulation steps (integrator kernels + force path, nothing but the time crosses PCIe)
    resident = None
    if world == 1:
        pr = dict(p)
        for k in ("vx", "vy", "vz", "U", "next_time", "mu"):
            pr[k] = p[k]
        ctx.set_particles(pr)
        ctx.integrator_init(2.0, 1e13, 1e13, 70.0, e0)               # the shipped Config.ini: fixed dt = 1e13 s => every particle active every step
        R = ctx.build_tree(); ctx.visual_density(R / 100000); ctx.gas_density(mh); ctx.forces(0.0, e0, THETA)
        ctx.integrator_assign_all()
        def res_step():
            t = ctx.step_begin()
            ctx.force_path(R / 100000, mh, t, e0, THETA)
            ctx.step_end()
        try:
            for _ in range(2):
                res_step()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            nres = min(20, max(3, args.steps // 2))
            for _ in range(nres):
                res_step()
            torch.cuda.synchronize()
            tr = (time.perf_counter() - t0) / nres
            resident = {"value": n / tr, "unit": "particles/s", "ms_per_step": tr * 1e3, "steps": nres,
                        "what": "full KDK simulation step (re-binning, kick, drift, tree, densities, forces, Ueuler, Hubble, kick) with state resident in HBM"}
        except Exception as e:  # noqa: BLE001
            # the shipped fixed dt = 1e13 s is far too long for the densest synthetic sets (64M merger): close encounters fling
            # particles out after a few steps and the rest of the system then sits deeper than 42 octree levels (AGB_ERR_DEPTH)
            resident = {"error": str(e), "what": "device-resident KDK loop with the shipped fixed time step"}

    if breakdown:
        torch.cuda.synchronize()
        ex = np.mean([e[0].elapsed_time(e[1]) for e, _ in bd_ev[args.warmup:args.warmup + args.steps]])
        pa = np.mean([e[1].elapsed_time(e[2]) for e, _ in bd_ev[args.warmup:args.warmup + args.steps]])
        ph = {k: float(np.mean([q[k] for _, q in bd_ev[args.warmup:args.warmup + args.steps]])) for k in bd_ev[0][1]}
        print("rank %d breakdown: exchange %.3f ms, path call %.3f ms, phases %s" % (rank, ex, pa, json.dumps(ph)), file=sys.stderr, flush=True)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (k_walk): CUDA-core FP pipe (no tensor cores: irregular, not a dense contraction).
    # Pair forces run in FP32 (packed FFMA2/FADD2) in the default mixed mode, traversal decisions in FP64; the denominator is
    # the FP32 FMA throughput measured on this GPU by agb_microbench (MEASURED_PEAKS.json has no CUDA-core figure).
    flop_per_interaction = 21.5                      # 10 flop per node visit x 1.15 + 10 flop per accepted pair
    walk_ms_avg = walk_total_ms / args.steps
    fp32_peak, fp64_peak = ctx.microbench(1), ctx.microbench(0)
    inter_rank = inter_all / world
    achieved = flop_per_interaction * inter_rank / (walk_ms_avg * 1e-3) / 1e12
    traffic = None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "walk_dram_traffic.json")))
        traffic = tr.get(name, {}).get("bytes_per_launch")
    except Exception:  # noqa: BLE001
        pass
    roofline = {"bound": "fp32", "kernel": "k_walk<mixed>", "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s", "frac": achieved / fp32_peak if fp32_peak else None,
                "traffic": traffic, "peak_source": "measured on this GPU: FP32 FMA chain microbenchmark (agb_microbench kind 1); FP64 chain = %.1f TFLOP/s" % fp64_peak,
                "algorithmic_flop_per_interaction": flop_per_interaction, "interactions_per_launch": inter_rank,
                "interactions_per_s": inter_all / (walk_ms_avg * 1e-3), "walk_ms": walk_ms_avg, "build_ms": float(np.mean(build_ms)),
                "note": "DRAM traffic of the walk is ~0.1 GB per launch (tree is L2 resident): not HBM bound"}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        hbm_peak, hbm_src = float(peaks["hbm_gbs"]), "MEASURED_PEAKS.json"
    except Exception:  # noqa: BLE001
        hbm_peak, hbm_src = 6650.0, "fallback (B200_PROFILING.md)"
    # build: key-gen 24+12 B, 8-pass sort of 12-byte items 2*8*12 B + 8*8 B histogram reads, permute ~100 B, links + upward ~190 B
    build_bytes = (36.0 + 256.0 + 100.0 + 190.0) * n
    bms = float(np.mean(build_ms)) * 1e-3
    roofline["hbm_build"] = {"bound": "hbm", "kernels": "k_keygen k_sort_* k_gather k_links k_upward ...", "achieved": build_bytes / bms / 1e9, "peak": hbm_peak, "unit": "GB/s",
                             "frac": build_bytes / bms / 1e9 / hbm_peak, "peak_source": hbm_src, "algorithmic_bytes_per_particle": build_bytes / n}

    # ---- CPU baseline on this box's host cores (bounded sample of the same workload)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        ns = sample_size_for_cpu(n, 2, budget_s=20.0)
        ps, _, _, _ = make_particles(pkg, name, ns)
        with tempfile.TemporaryDirectory() as d:
            secs, kind, cores, rows = cpu_reference_run(ps, e0, mh, 2, d)
        cpu = {"value": ns / secs[-1], "unit": "particles/s", "cores": cores, "kind": kind,
               "sample": "%d of %d particles of %s, 2nd of 2 steps, phases build+visual+gas_density+forces = %.3f s" % (ns, n, name, secs[-1])}

    line = {
        "metric": METRIC, "value": value, "unit": "particles/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64 decisions and accumulation, f32 pair forces (mixed mode)", "data": "synthetic",
        "config": {"workload": name, "description": desc, "n_particles": n, "theta": THETA, "e0": e0, "massInH": mh, "all_active": True, "precision": "mixed",
                   "l2": "256 MiB buffer written between timed steps (L2 flush); working set %.0f MB" % (n * 330 / 1e6),
                   "parallelism": "replicated tree, tree-ordered target slices, 1 coalesced NCCL all-gather group/step (in place)" if world > 1 else "single GPU"},
        "e2e": e2e, "resident_sim_step": resident, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
        "interactions_per_s": inter_all / (walk_ms_avg * 1e-3), "wall_s_timed_region": t_wall,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="agb200", choices=["agb200", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    import __graft_entry__ as ge
    pkg = ge.load_package()
    if args.impl == "reference":
        run_reference_arm(args, pkg)
    else:
        run_gpu_arm(args, pkg)


if __name__ == "__main__":
    main()
