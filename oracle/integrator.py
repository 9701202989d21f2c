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


def run_steps(p, forces, e0, eta, min_ts, max_ts, H0, nsteps):
    st = {k: np.array(v, copy=True) for k, v in p.items()}
    n = len(st["x"])
    for k in ("ax", "ay", "az", "dUdt", "h", "vis"):
        st.setdefault(k, np.zeros(n))
    st["timeStep"] = np.zeros(n)
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
