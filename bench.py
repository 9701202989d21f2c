#!/usr/bin/env python
"""bench.py — headline benchmark of the force hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

A step = one pass of the path over one synthetic particle set: build_tree -> visual_density ->
gas_density -> forces (the reference's Simulation::run step, Simulation.cpp:276-285), all particles
active.  Default workload at every N: BASELINE.json's north-star configuration C3, the gas-rich disk
galaxy with 16M particles (gravity + SPH); `--workload plummer1m` is C1 (gravity only).  `value` = particle-updates/s with the particles already resident in HBM; `e2e` = the same
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


DEFAULT_WORKLOAD = "gas16m"


def workload_config(pkg, name):
    """The `config` object of the JSON line: identical (keys and values) in the GPU arm and in --impl reference."""
    gen, n, e0, nb, desc = WORKLOADS[name]
    # massInH = nb gas-particle masses of the FULL-size set (equal-mass generators: M_tot / n per particle)
    mtot = {"plummer": 1e11, "disk": 1e12, "gasdisk": 1e12, "merger": 2e12}[gen] * pkg.ics.MSUN
    mh = float(nb * (mtot / n)) if nb else 1e40
    return {"workload": name, "description": desc, "n_particles": n, "theta": THETA, "e0": e0, "massInH": mh, "all_active": True}


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
    """The reference's own CPU implementation of the path (unmodified sources compiled in place: oracle/_ref/ag_ref_omp, all host
    threads) on the GPU arm's config.  Each step is a bounded sample of the workload: the same generator at a particle count
    sized so that the whole --steps/--warmup run ends within a few minutes.  The reference's throughput FALLS with the
    particle count (deeper tree, more interactions per target), so the smaller sample favours the CPU arm."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    name = args.workload or DEFAULT_WORKLOAD
    cfg = workload_config(pkg, name)
    n_full = cfg["n_particles"]
    total = args.steps + args.warmup
    n = sample_size_for_cpu(n_full, total, budget_s=150.0)
    p, e0, _, desc = make_particles(pkg, name, n)
    nb = WORKLOADS[name][3]
    mh = cfg["massInH"] * (n_full / n) if nb else cfg["massInH"]           # the same number of gas-particle masses in the sample
    with tempfile.TemporaryDirectory() as d:
        secs, kind, cores, rows = cpu_reference_run(p, e0, mh, total, d)
    timed = secs[args.warmup:]
    t = float(np.mean(timed))
    value = n / t
    sample = ("%d of %d particles (same generator and seed at the smaller count; the reference's per-particle cost grows with N, so this favours the CPU), "
              "%d timed steps, phases build+visual+gas_density+forces" % (n, n_full, len(timed))) if n < n_full else \
             ("all %d particles, %d timed steps, phases build+visual+gas_density+forces" % (n_full, len(timed)))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "particles/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": value, "unit": "particles/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "particles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "phases_s": rows[-1],
    }
    print(json.dumps(line))


# ------------------------------------------------------------------ GPU arm
RESULT_COLS_GRAVITY = ("ax", "ay", "az", "visualDensity")
RESULT_COLS_GAS = ("dUdt", "h", "rho", "P", "T")


def bind_near_gpu(torch, local):
    """Several ranks on one node: keep this process (and the pinned host buffers it is about to allocate, first touch) on the NUMA
    node its GPU hangs off, so that N ranks do not push their host<->device traffic across the socket interconnect.  Best effort:
    returns the node, or None when the topology is not visible (containers) or the node's CPUs are not allowed."""
    try:
        pr = torch.cuda.get_device_properties(local)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        use = os.sched_getaffinity(0) & cpus
        if not use:
            return None
        os.sched_setaffinity(0, use)
        return node
    except Exception:  # noqa: BLE001
        return None


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
    numa_node = bind_near_gpu(torch, local) if world > 1 and os.environ.get("AGB_BENCH_NO_AFFINITY") != "1" else None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    name = args.workload or DEFAULT_WORKLOAD
    cfg = workload_config(pkg, name)
    p, e0, _, desc = make_particles(pkg, name)
    mh = cfg["massInH"]
    n = len(p["x"])
    any_gas = bool((p["type"] == 2).any())
    ctx = pkg.Context(local, 8)
    if args.extended:
        # SURVEY.md §8(f)-3, reported on its own lines: quadrupole walk with spline softening + per-particle-h neighbour SPH (FP64; not the
        # reference's algorithm, so there is no reference arm and no parity claim for this line)
        ctx.set_option(pkg.capi.AGB_OPT_EXTENDED, 1)
        args.no_fp64 = True
    stream = torch.cuda.ExternalStream(ctx.stream(), device=dev)
    out_cols = RESULT_COLS_GRAVITY + (RESULT_COLS_GAS if any_gas else ())      # what a step hands back, for any number of GPUs

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

    # Staged exchange (equal shards): positions / masses / types first, then next_time, then velocities / U / mu, each group as one
    # coalesced in-place all-gather with an event behind it.  agb_set_particles_staged reads a group only after its event, so the
    # tree build, the densities and the gravity walk overlap the exchange (and, end to end, the upload) of the later groups.
    staged = world > 1 and isinstance(packed, pkg.shard.InPlaceGather)
    groups = [[k for k in g if k in shard] for g in (("x", "y", "z", "mass", "type"), ("next_time",), ("vx", "vy", "vz", "U", "mu"))]
    keep_events = []

    breakdown = os.environ.get("AGB_BENCH_BREAKDOWN") == "1"      # development: per-step device time of the exchange vs the path
    bd_ev = []

    def step(upload=None):
        if breakdown:
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            e[0].record(torch.cuda.current_stream())
        ptrs = {k: full[k].data_ptr() for k in full}
        if staged and upload is None:
            # shards already on the device — two groups: what extent, keys and sort read, then everything else behind one event (no
            # late group: the build joins it before the gather), so the exchange of the second group hides behind the first 2 ms of
            # the build.  (Three groups with the late part of the build cost 1.4 ms of extra kernels on C3, more than the exchange
            # they hide: 25.1 vs 24.6 ms at 2 GPUs; with the upload in the loop, below, they hide 10+ ms.)
            del keep_events[:]
            for grp in (groups[0], groups[1] + groups[2]):
                if grp:
                    packed.gather_fields(grp)
                ev_ = torch.cuda.Event()
                ev_.record(torch.cuda.current_stream())
                keep_events.append(ev_)
            if breakdown:
                e[1].record(torch.cuda.current_stream())
            ctx.set_particles_device(ptrs, n, events=[keep_events[0].cuda_event, keep_events[1].cuda_event, keep_events[1].cuda_event])
        elif staged:
            evs = []
            del keep_events[:]
            for grp in groups:
                if not grp:
                    evs.append(0)
                    continue
                for k in grp:
                    shard[k].copy_(upload[k], non_blocking=True)
                packed.gather_fields(grp)
                ev_ = torch.cuda.Event()
                ev_.record(torch.cuda.current_stream())
                keep_events.append(ev_)
                evs.append(ev_.cuda_event)
            if breakdown:
                e[1].record(torch.cuda.current_stream())
            ctx.set_particles_device(ptrs, n, events=evs)
        else:
            if upload is not None:
                for k in upload:
                    shard[k].copy_(upload[k], non_blocking=True)
            gather()
            if breakdown:
                e[1].record(torch.cuda.current_stream())
            if world > 1:
                stream.wait_stream(torch.cuda.current_stream())
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
    multi_gpu_check = None
    if world > 1:                                                              # the exchange really delivered every rank's shard (checked once, untimed)
        for k in full:
            if not torch.equal(full[k], torch.from_numpy(np.ascontiguousarray(p[k])).to(dev)):
                raise SystemExit("all-gather of %s does not reproduce the particle set" % k)
    ctx.set_particles_device({k: full[k].data_ptr() for k in full}, n)
    vis_radius = ctx.build_tree() / 100000                                     # Simulation.cpp:123-126
    if world > 1:
        # Untimed, once: the compact results of this rank's slice of an N-way sharded walk equal, bit for bit, the same
        # particles' results of a whole (1-GPU style) walk of the same tree.
        ctx.visual_density(vis_radius); ctx.gas_density(mh)
        ctx.forces(0.0, e0, THETA, rank, world)
        mine = ctx.slice_results(rank, world, names=out_cols)
        mine = {k: v.copy() for k, v in mine.items()}
        mine_check = mine
        ctx.set_particles_device({k: full[k].data_ptr() for k in full}, n)     # fresh carried state (dU/dt accumulates across calls)
        ctx.build_tree(); ctx.visual_density(vis_radius); ctx.gas_density(mh)
        ctx.forces(0.0, e0, THETA, 0, 1)
        whole = ctx.results(names=out_cols)
        same = all(np.array_equal(whole[k][mine["index"]], mine[k]) for k in out_cols)
        flag = torch.tensor([1.0 if same else 0.0, float(len(mine["index"]))], dtype=torch.float64, device=dev)
        mn = flag.clone(); dist.all_reduce(mn, op=dist.ReduceOp.MIN)
        sm_ = flag.clone(); dist.all_reduce(sm_, op=dist.ReduceOp.SUM)
        if float(mn[0]) != 1.0 or int(sm_[1]) != n:
            raise SystemExit("sharded walk differs from the whole walk (or the slices do not cover every particle once)")
        multi_gpu_check = "slices of the %d-way sharded walk cover all %d particles once and equal the whole walk bit for bit (%s)" % (world, n, ", ".join(out_cols))
        # every rank hands back its slice only: the density passes produce their per-particle outputs for that slice only (the SPH
        # terms of a target use its own h, rho, P; the end-to-end leg below checks the delivered columns against the bits above)
        if os.environ.get("AGB_BENCH_FULL_DENSITIES") != "1":
            ctx.set_option(pkg.capi.AGB_OPT_SLICE_DENSITIES, 1)
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
    kms = {}
    inter = sph_pairs = 0
    barrier()
    t_wall0 = time.perf_counter()
    for i in range(args.steps):
        flush.fill_(i & 0xff)                      # L2 flush between timed iterations (outside the timed interval)
        barrier()
        ev[i][0].record(stream if world == 1 else torch.cuda.current_stream())
        step()
        ev[i][1].record(stream)
        ev[i][1].synchronize()
        for k, v in list(ctx.phase_ms().items()) + list(ctx.kernel_ms().items()):
            kms.setdefault(k, []).append(v)
        cnt = ctx.counters()
        inter, sph_pairs = cnt["interactions"], cnt["sph_interactions"]
    barrier()
    t_wall = time.perf_counter() - t_wall0
    step_ms = [a.elapsed_time(b) for a, b in ev]
    launches = ctx.launch_count() - launches0
    clocks = sampler.stop(t_wall0, t_wall0 + t_wall) if rank == 0 else None
    kavg = {k: float(np.mean(v)) for k, v in kms.items()}
    tot = torch.tensor([sum(step_ms), float(inter), kavg["k_walk"], float(sph_pairs), kavg["k_sph"]], dtype=torch.float64, device=dev)
    if world > 1:
        mx = tot.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = tot.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        total_ms, walk_ms_avg, sph_ms_avg, inter_all, sph_all = float(mx[0]), float(mx[2]), float(mx[4]), float(sm[1]), float(sm[3])
    else:
        total_ms, walk_ms_avg, sph_ms_avg, inter_all, sph_all = float(tot[0]), float(tot[2]), float(tot[4]), float(inter), float(sph_pairs)
    ms_per_step = total_ms / args.steps
    value = n / (ms_per_step * 1e-3)
    divergence = {k: cnt[k] for k in ("edge_dropped", "gas_ties_unresolved", "mac_exact_fallbacks", "walk_stack_spills", "n_outliers", "max_depth", "gas_orphans")}

    # ---- the same steps with FP64 pair arithmetic throughout (AGB_OPT_PRECISION = 0): the reference's own arithmetic type
    fp64 = None
    if not args.no_fp64:
        ctx.set_option(pkg.capi.AGB_OPT_PRECISION, 0)
        nf = max(3, args.steps // 4)
        for _ in range(2):
            step()
        barrier()
        evf = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(nf)]
        wf = []
        for i in range(nf):
            flush.fill_(i & 0xff)
            barrier()
            evf[i][0].record(stream if world == 1 else torch.cuda.current_stream())
            step()
            evf[i][1].record(stream)
            evf[i][1].synchronize()
            wf.append(ctx.kernel_ms()["k_walk"])
        barrier()
        tf = torch.tensor([sum(a.elapsed_time(b) for a, b in evf) / nf, float(np.mean(wf))], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tf, op=dist.ReduceOp.MAX)
        fp64 = {"ms_per_step": float(tf[0]), "value": n / (float(tf[0]) * 1e-3), "unit": "particles/s", "steps": nf, "k_walk_ms": float(tf[1]),
                "what": "same workload with FP64 pair forces and in-walk FP64 SPH (agrees with the reference to ~1e-14)"}
        ctx.set_option(pkg.capi.AGB_OPT_PRECISION, 1)

    # ---- end-to-end through the Tree API with pinned host arrays (N GPUs: every rank uploads its shard, results of its slice come back)
    e2e = None
    if world == 1:
        host = {}
        for k in f8 + ["type"]:
            t = torch.from_numpy(np.ascontiguousarray(p[k])).pin_memory()
            host[k] = t.numpy()
        pinned_out = {k: torch.empty(n, dtype=torch.float64).pin_memory() for k in out_cols}

        out_np = {k: v.numpy() for k, v in pinned_out.items()}
        # caller-order delivery: density outputs leave while the walk runs; acc / dUdt follow in results_into. (Measured alternative on C3:
        # the N>1 form — agb_bind_slice_results(0, 1), tree-order pieces behind the walk plus the index — 62.0 ms against 61.5 ms for this
        # one, and it leaves the caller a permutation to do; profiles/README.md.)
        ctx.bind_results(out_np)

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
        e2e = {"value": n / te, "unit": "particles/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "ms_per_step": te * 1e3,
               "what": "agb_set_particles (pinned host arrays) -> agb_force_path -> agb_get_results (%s) per step" % ", ".join(out_cols)}
        ctx.bind_results(None)
    else:
        # multi-GPU e2e: host shard -> device shard (H2D), gather, step, then every rank brings back the compact results of
        # ITS slice of the targets — the same columns as the 1-GPU run plus the caller-order index (agb_get_slice_results_all)
        host = {k: torch.from_numpy(np.ascontiguousarray(p[k][lo:hi])).pin_memory() for k in f8 + ["type"]}
        cap = n // world + 1024
        out_t = {k: torch.empty(cap, dtype=torch.float64).pin_memory() for k in out_cols}
        out_t["index"] = torch.empty(cap, dtype=torch.int32).pin_memory()
        out_np = {k: v.numpy() for k, v in out_t.items()}
        out_np["index"] = out_np["index"].view(np.uint32)

        cnt_mine = ctx.slice_count(rank, world)
        ctx.bind_slice_results(rank, world, out_np)       # agb_force_path delivers this rank's slice itself: densities during the walk, acc / dU/dt after it

        def e2e_step():
            step(upload=host)
            if breakdown and rank == 0:
                e_, _ = bd_ev[-1]
                torch.cuda.synchronize()
                print("rank 0 e2e step: uploads + exchanges done %.2f ms after the step began, path done %.2f ms" % (e_[0].elapsed_time(e_[1]), e_[0].elapsed_time(e_[2])), file=sys.stderr, flush=True)
            return {k: v[:cnt_mine] for k, v in out_np.items()}
        for v in out_np.values():
            v.fill(0)
        r = e2e_step()
        # the staged, overlapped step returns the bits of the untimed call-by-call walk of this slice
        if not (np.array_equal(r["index"], mine_check["index"]) and all(np.array_equal(r[k], mine_check[k]) for k in out_cols)):
            raise SystemExit("end-to-end step differs from the checked slice results")
        mine = len(r["index"]) * (4 + 8 * len(out_cols))
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
        ctx.bind_slice_results(rank, world, None)
        e2e = {"value": n / te, "unit": "particles/s", "h2d_bytes_per_step": int(sum(v.numel() * v.element_size() for v in host.values()) * world),
               "d2h_bytes_per_step": int(by[1]), "ms_per_step": te * 1e3, "host_numa_node_rank0": numa_node,
               "what": "per rank: H2D of its particle shard and NCCL all-gather in three groups (positions+mass+type | next_time | velocities, U, mu) overlapped with build, densities and walk of its target slice, D2H of that slice's (index, %s)" % ", ".join(out_cols)}

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
            # 2 + 3 steps only: with the shipped fixed step (1e13 s) the gas of C3 / C4 starts to blow up after ~5 steps (a few particles
            # are ejected at >> c, the root cube grows by 2^20 within a few steps); the path follows that (three-word keys, FP64
            # pair arithmetic outside the FP32 law's range) but those steps are not the workload this line describes
            nres = 3
            for _ in range(nres):
                res_step()
            torch.cuda.synchronize()
            tr = (time.perf_counter() - t0) / nres
            resident = {"value": n / tr, "unit": "particles/s", "ms_per_step": tr * 1e3, "steps": nres,
                        "what": "full KDK simulation step (re-binning, kick, drift, tree, densities, forces, Ueuler, Hubble, kick) with state resident in HBM"}
        except Exception as e:  # noqa: BLE001
            resident = {"error": str(e), "what": "device-resident KDK loop with the shipped fixed time step"}

    if breakdown:
        torch.cuda.synchronize()
        ex = np.mean([e[0].elapsed_time(e[1]) for e, _ in bd_ev[args.warmup:args.warmup + args.steps]])
        pa = np.mean([e[1].elapsed_time(e[2]) for e, _ in bd_ev[args.warmup:args.warmup + args.steps]])
        ph = {k: float(np.mean([q[k] for _, q in bd_ev[args.warmup:args.warmup + args.steps]])) for k in bd_ev[0][1]}
        print("rank %d breakdown: exchange %.3f ms, path call %.3f ms, phases %s kernels %s" % (rank, ex, pa, json.dumps(ph), json.dumps(kavg)), file=sys.stderr, flush=True)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (k_walk): CUDA-core FP pipe (no tensor cores: irregular, not a dense contraction).
    # Pair forces run in FP32 (packed FFMA2/FADD2) in the default mixed mode, traversal decisions in FP64; the denominator is
    # the FP32 FMA throughput measured on this GPU by agb_microbench (MEASURED_PEAKS.json has no CUDA-core figure).
    flop_per_interaction = 21.5                      # 10 flop per node visit x 1.15 + 10 flop per accepted pair
    flop_per_sph_pair = 55.0
    fp32_peak, fp64_peak = ctx.microbench(1), ctx.microbench(0)
    inter_rank = inter_all / world
    achieved = flop_per_interaction * inter_rank / (walk_ms_avg * 1e-3) / 1e12
    traffic = traffic_src = None
    kname = "k_walk<COUNT=0,SPH=%d,MIXED=1>" % (1 if any_gas else 0)
    if args.extended:
        # monopole with spline softening ~23 flop per pair, + ~37 for the quadrupole term of a node source
        kname = "k_walk_ext (kernel_ms: k_far = quadrupole upward pass, k_walk = k_walk_ext, k_sph = k_ext_sph_force)"
        flop_per_interaction = (23.0 * cnt["leaf_interactions"] + 60.0 * cnt["node_interactions"]) / max(1, cnt["interactions"])
        fp32_peak = fp64_peak
        achieved = flop_per_interaction * inter_rank / (walk_ms_avg * 1e-3) / 1e12
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "walk_dram_traffic.json")))
        ent = tr.get(name + ("_extended" if args.extended else ""), {})
        traffic = ent.get("bytes_per_launch")
        traffic_src = ent.get("source")
    except Exception:  # noqa: BLE001
        pass
    roofline = {"bound": "fp64" if args.extended else "fp32", "kernel": kname, "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s", "frac": achieved / fp32_peak if fp32_peak else None,
                "traffic": traffic, "traffic_source": traffic_src,
                "peak_source": "measured on this GPU: FP32 FMA chain microbenchmark (agb_microbench kind 1); FP64 chain = %.1f TFLOP/s" % fp64_peak,
                "algorithmic_flop_per_interaction": flop_per_interaction, "interactions_per_launch": inter_rank,
                "interactions_per_s": inter_all / (walk_ms_avg * 1e-3), "k_walk_ms": walk_ms_avg, "kernel_ms": kavg,
                "note": "the walk's DRAM traffic is a few %% of HBM peak (tree L2 resident): bounded by the CUDA-core issue rate, not by HBM"}
    if any_gas and sph_ms_avg > 0:
        a = flop_per_sph_pair * (sph_all / world) / (sph_ms_avg * 1e-3) / 1e12
        roofline["k_sph"] = {"bound": "fp32", "achieved": a, "peak": fp32_peak, "unit": "TFLOP/s", "frac": a / fp32_peak if fp32_peak else None,
                             "algorithmic_flop_per_pair": flop_per_sph_pair, "pairs_per_launch": sph_all / world, "k_sph_ms": sph_ms_avg}
    if fp64 is not None:
        a = flop_per_interaction * inter_rank / (fp64["k_walk_ms"] * 1e-3) / 1e12
        fp64["roofline"] = {"bound": "fp64", "kernel": "k_walk<COUNT=0,SPH=%d,MIXED=0>" % (1 if any_gas else 0), "achieved": a, "peak": fp64_peak, "unit": "TFLOP/s",
                            "frac": a / fp64_peak if fp64_peak else None, "peak_source": "FP64 FMA chain microbenchmark on this GPU"}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        hbm_peak, hbm_src = float(peaks["hbm_gbs"]), "MEASURED_PEAKS.json"
    except Exception:  # noqa: BLE001
        hbm_peak, hbm_src = 6650.0, "fallback (B200_PROFILING.md)"
    # build: key-gen 24+12 B, 8-pass sort of 12-byte items 2*8*12 B + 8*8 B histogram reads, permute ~100 B, links + upward ~190 B
    build_bytes = (36.0 + 256.0 + 100.0 + 190.0) * n
    bms = kavg["build"] * 1e-3
    roofline["hbm_build"] = {"bound": "hbm", "kernels": "k_keygen k_sort_* k_gather k_links k_upward ...", "achieved": build_bytes / bms / 1e9, "peak": hbm_peak, "unit": "GB/s",
                             "frac": build_bytes / bms / 1e9 / hbm_peak, "peak_source": hbm_src, "algorithmic_bytes_per_particle": build_bytes / n, "build_ms": kavg["build"]}

    # ---- CPU baseline on this box's host cores (bounded sample of the same workload)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        ns = sample_size_for_cpu(n, 2, budget_s=20.0)
        ps, _, _, _ = make_particles(pkg, name, ns)
        mhs = mh * (n / ns) if WORKLOADS[name][3] else mh
        with tempfile.TemporaryDirectory() as d:
            secs, kind, cores, rows = cpu_reference_run(ps, e0, mhs, 2, d)
        cpu = {"value": ns / secs[-1], "unit": "particles/s", "cores": cores, "kind": kind,
               "sample": "%d of %d particles of %s (same generator at the smaller count: favours the CPU), 2nd of 2 steps, phases build+visual+gas_density+forces = %.3f s" % (ns, n, name, secs[-1])}

    line = {
        "metric": METRIC, "value": value, "unit": "particles/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64" if args.extended else "f64 decisions and accumulation, f32 pair forces (mixed mode)", "data": "synthetic",
        "config": cfg,
        "run": {"precision": "extended-accuracy mode (AGB_OPT_EXTENDED): quadrupoles, spline softening, width/d < theta per 32-target group, per-particle-h SPH; FP64; parity unpinned by the reference" if args.extended else "mixed", "l2": "256 MiB buffer written between timed steps (L2 flush); working set %.0f MB" % (n * 330 / 1e6),
                "parallelism": ("replicated tree, tree-ordered target slices, density outputs for the rank's slice only, in-place NCCL all-gathers in 2 coalesced groups per step, the second behind extent / keys / sort (e2e: 3 groups, overlapped with build, densities and gravity walk) through agb_set_particles_staged" if staged else
                                "replicated tree, tree-ordered target slices, packed NCCL all-gather per step") if world > 1 else "single GPU",
                "result_columns": list(out_cols)},
        "e2e": e2e, "fp64": fp64, "resident_sim_step": resident, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
        "interactions_per_s": inter_all / (walk_ms_avg * 1e-3), "sph_pairs_per_step": sph_all, "sph_records_per_step_rank0": cnt.get("sph_records"), "divergence_counters": divergence, "multi_gpu_check": multi_gpu_check,
        "wall_s_timed_region": t_wall,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="agb200", choices=["agb200", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-fp64", action="store_true", help="skip the FP64-arithmetic leg")
    ap.add_argument("--extended", action="store_true", help="extended-accuracy mode (quadrupoles, spline softening, neighbour-loop SPH): its own line, no reference arm")
    args = ap.parse_args()
    import __graft_entry__ as ge
    pkg = ge.load_package()
    if args.impl == "reference":
        run_reference_arm(args, pkg)
    else:
        run_gpu_arm(args, pkg)


if __name__ == "__main__":
    main()
