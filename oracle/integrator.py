"""oracle/integrator.py — numpy restatement of the reference's driver loop around the force path.  TEST INFRASTRUCTURE ONLY.

Follows Simulation::init (force part, Simulation.cpp:101-139) and Simulation::run (:166-345) with the
reference's TimeIntegration (TimeIntegration.cpp:10-41), operation order preserved so that, with the CPU
oracle as the force provider, the trajectory is bit-identical to oracle/_ref/ag_ref `steps`.
`forces(state, globalTime)` must return a dict with ax ay az dUdt h rho P T vis for all particles."""
import math

import numpy as np

GAMMA, KB, PRTN = 5.0 / 3.0, 1.38064852e-23, 1.6726219e-27
KMS, MPC = 1.0e3, 3.08567758149137e22


def _assign(st, sel, gt, eta, e0, min_ts, max_ts):
    a = np.sqrt(st["ax"][sel] * st["ax"][sel] + st["ay"][sel] * st["ay"][sel] + st["az"][sel] * st["az"][sel])
    ts = np.full(a.shape, min_ts)
    pos = a > 0
    with np.errstate(divide="ignore", invalid="ignore"):
        t = eta * np.sqrt(e0 / a[pos])
    t = np.minimum(np.maximum(t, min_ts), max_ts)                      # std::clamp
    _, ex = np.frexp(t)                                                # floor(log2(t)) == exponent - 1
    t = np.maximum(np.ldexp(1.0, ex - 1), min_ts)
    ts[pos] = t
    st["timeStep"][sel] = ts
    st["next_time"][sel] = gt + ts


def u01(seed, idx, gt):
    """The library's agb_u01 (agb_internal.cuh): splitmix64 of (seed, particle, time bits) -> [0, 1)."""
    M = (1 << 64) - 1
    tb = int(np.float64(gt).view(np.uint64))
    out = np.empty(len(idx))
    for k, i in enumerate(idx):
        z = (seed + int(i) * 0x9E3779B97F4A7C15 + tb * 0xD1B54A32D192ED03) & M
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M
        z ^= z >> 31
        out[k] = (z >> 11) * (1.0 / 9007199254740992.0)
    return out


def run_steps(p, forces, e0, eta, min_ts, max_ts, H0, nsteps, cooling=False, sf_seed=0):
    """cooling / sf_seed != 0: the sub-grid hooks the reference's loop calls for active gas before Ueuler (Simulation.cpp:311-320,
    commented out there; Cooling.cpp:6-25, SFR.cpp:12-34 with a counter-based deviate in place of rand())."""
    st = {k: np.array(v, copy=True) for k, v in p.items()}
    n = len(st["x"])
    for k in ("ax", "ay", "az", "dUdt", "h", "vis"):
        st.setdefault(k, np.zeros(n))
    st["timeStep"] = np.zeros(n)
    st["sfr"] = np.zeros(n)
    gas = st["type"] == 2
    st["T"][gas] = (GAMMA - 1.0) * st["U"][gas] * PRTN * st["mu"][gas] / KB          # Simulation.cpp:108-112
    gt = 0.0
    st.update(forces(st, gt))                                                          # Simulation.cpp:120-139
    st["next_time"][:] = 0.0
    _assign(st, np.ones(n, bool), gt, eta, e0, min_ts, max_ts)
    H0SI = (H0 * KMS) / MPC
    for _ in range(nsteps):
        due = gt >= st["next_time"]
        if due.any():
            _assign(st, due, gt, eta, e0, min_ts, max_ts)
        gt = float(st["next_time"].min())
        act = st["next_time"] == gt
        dt = st["timeStep"]
        ok = ~(np.isnan(st["ax"]) | np.isnan(st["ay"]) | np.isnan(st["az"])) & act
        for v, a in (("vx", "ax"), ("vy", "ay"), ("vz", "az")):                        # Kick: v + acc * dt / 2
            st[v][ok] = st[v][ok] + st[a][ok] * dt[ok] / 2
        for x, v in (("x", "vx"), ("y", "vy"), ("z", "vz")):                           # Drift
            st[x][act] = st[x][act] + st[v][act] * dt[act]
        st.update(forces(st, gt))
        gas = st["type"] == 2
        ga = act & gas
        if cooling:                                                                    # Cooling.cpp:6-25
            rate = 1.42e-27 * 1.1 * np.sqrt(st["T"][ga]) * 1e6 * 1e6 * 1e-7
            okc = (rate > 0) & (st["rho"][ga] > 0)
            idx = np.flatnonzero(ga)[okc]
            st["dUdt"][idx] = st["dUdt"][idx] - rate[okc] / st["rho"][idx]
        if sf_seed:                                                                    # SFR.cpp:12-34
            cand = np.flatnonzero(ga & (st["rho"] > 1e-22) & (st["T"] < 1e4))
            if len(cand):
                pr = np.array([1 - math.exp(-0.1 * d / 1e15) for d in dt[cand]])
                st["sfr"][cand] = pr
                born = cand[u01(sf_seed, cand, gt) < pr]
                st["type"][born] = 1
                st["U"][born] = 0.0
            gas = st["type"] == 2
            ga = act & gas
        good = ga & ~np.isnan(st["dUdt"])
        st["U"][good] = st["U"][good] + st["dUdt"][good] * dt[good]                    # Ueuler
        st["dUdt"][ga] = 0
        scale = np.ones(n)
        for d in np.unique(dt[act]):
            scale[act & (dt == d)] = math.exp(H0SI * d)                                # libm exp, like the reference
        for x in ("x", "y", "z"):
            st[x][act] = st[x][act] * scale[act]
        ok = ~(np.isnan(st["ax"]) | np.isnan(st["ay"]) | np.isnan(st["az"])) & act
        for v, a in (("vx", "ax"), ("vy", "ay"), ("vz", "az")):
            st[v][ok] = st[v][ok] + st[a][ok] * dt[ok] / 2
        st["next_time"][act] = st["next_time"][act] + dt[act]
    st["globalTime"] = gt
    return st


def oracle_forces(theta, e0, massInH, cores):
    from . import oracle

    def f(st, gt):
        o = oracle.run(st, theta, e0, massInH, gt, cores, nodes=False, counters=False)
        return {k: o[k] for k in ("ax", "ay", "az", "dUdt", "h", "rho", "P", "T", "vis")}
    return f
