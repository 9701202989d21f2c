"""Synthetic initial conditions for the benchmark / parity configurations of SURVEY.md §8(d).

All SI units (the reference path works in SI: |x| ~ 1e20..1e25 m, m ~ 1e35 kg).  The generator is
numpy PCG64 with a fixed seed; particle order is generator order (no shuffle) so the CPU checker and the
GPU path see identical arrays.  Particle fields follow the reference's record
(Physics/Particle.h:18-57): type 1 = star, 2 = gas, 3 = dark matter; mu defaults to 0.58.
"""
import numpy as np

KPC = 3.08567758149137e19      # Math/Units.h
MSUN = 1.98847e30
G = 6.67430e-11


def _empty(n):
    p = {k: np.zeros(n) for k in ("x", "y", "z", "vx", "vy", "vz", "mass", "U", "next_time", "rho", "P", "T")}
    p["mu"] = np.full(n, 0.58)
    p["type"] = np.ones(n, dtype=np.uint8)
    return p


def _iso(rng, n):
    c = rng.uniform(-1.0, 1.0, n)
    ph = rng.uniform(0.0, 2.0 * np.pi, n)
    s = np.sqrt(1.0 - c * c)
    return s * np.cos(ph), s * np.sin(ph), c


def plummer(n, seed=1234, a=10 * KPC, mtot=1e11 * MSUN, gas_fraction=0.0, u_gas=1e9):
    """C1: Plummer sphere, a = 10 kpc, M = 1e11 Msun, equal masses, r = a / sqrt(X^(-2/3) - 1)."""
    rng = np.random.default_rng(seed)
    p = _empty(n)
    X = rng.uniform(1e-12, 1.0, n)
    r = a / np.sqrt(X ** (-2.0 / 3.0) - 1.0)
    ux, uy, uz = _iso(rng, n)
    p["x"], p["y"], p["z"] = r * ux, r * uy, r * uz
    # isotropic velocities with the local 1-D dispersion sigma^2 = G M / (6 sqrt(r^2 + a^2))
    sig = np.sqrt(G * mtot / (6.0 * np.sqrt(r * r + a * a)))
    p["vx"], p["vy"], p["vz"] = (sig * rng.standard_normal(n) for _ in range(3))
    p["mass"][:] = mtot / max(n, 1)
    if gas_fraction > 0:
        gas = rng.uniform(0, 1, n) < gas_fraction
        p["type"][gas] = 2
        p["U"][gas] = u_gas
    return p


def _hernquist_r(rng, n, a, rmax=None):
    # M(<r)/M = r^2/(r+a)^2  =>  r = a sqrt(X) / (1 - sqrt(X)); truncated at rmax by limiting X
    xmax = 1.0 if rmax is None else (rmax / (rmax + a)) ** 2
    s = np.sqrt(rng.uniform(0.0, xmax, n))
    return a * s / np.maximum(1.0 - s, 1e-12)


def disk_galaxy(n, seed=1234, gas_disk_fraction=0.25, mtot=1e12 * MSUN, u_gas=1e9,
                centre=(0.0, 0.0, 0.0), bulk_velocity=(0.0, 0.0, 0.0), rotation=None):
    """C2/C3: 50 % Hernquist halo (a = 30 kpc, r <= 300 kpc, type 3), 40 % exponential disk
    (R_d = 3 kpc, logistic z with z0 = 0.3 kpc), 10 % Hernquist bulge (a = 0.5 kpc); a fraction of the
    disk particles is gas (type 2, U = 1e9 J/kg); equal masses."""
    rng = np.random.default_rng(seed)
    p = _empty(n)
    nh = n // 2
    nd = (n * 4) // 10
    nb = n - nh - nd
    # halo
    r = _hernquist_r(rng, nh, 30 * KPC, 300 * KPC)
    ux, uy, uz = _iso(rng, nh)
    x = [r * ux]; y = [r * uy]; z = [r * uz]
    typ = [np.full(nh, 3, np.uint8)]
    # disk: surface density ~ exp(-R/Rd)  =>  R from the Gamma(2) distribution
    R = 3 * KPC * rng.gamma(2.0, 1.0, nd)
    ph = rng.uniform(0, 2 * np.pi, nd)
    zz = 0.3 * KPC * rng.logistic(0.0, 1.0, nd)
    x.append(R * np.cos(ph)); y.append(R * np.sin(ph)); z.append(zz)
    tdisk = np.ones(nd, np.uint8)
    tdisk[rng.uniform(0, 1, nd) < gas_disk_fraction] = 2
    typ.append(tdisk)
    # bulge
    r = _hernquist_r(rng, nb, 0.5 * KPC, 30 * KPC)
    ux, uy, uz = _iso(rng, nb)
    x.append(r * ux); y.append(r * uy); z.append(r * uz)
    typ.append(np.ones(nb, np.uint8))
    p["x"], p["y"], p["z"] = np.concatenate(x), np.concatenate(y), np.concatenate(z)
    p["type"] = np.concatenate(typ)
    p["mass"][:] = mtot / n
    # velocities: circular speed of the enclosed Hernquist halo mass in the disk plane + dispersion
    rr = np.sqrt(p["x"] ** 2 + p["y"] ** 2 + p["z"] ** 2) + 1e-3 * KPC
    menc = 0.5 * mtot * rr * rr / (rr + 30 * KPC) ** 2 + 0.5 * mtot * np.minimum(rr / (10 * KPC), 1.0)
    vc = np.sqrt(G * menc / rr)
    sig = 0.3 * vc
    p["vx"], p["vy"], p["vz"] = (sig * rng.standard_normal(n) for _ in range(3))
    d0, d1 = nh, nh + nd
    Rxy = np.sqrt(p["x"][d0:d1] ** 2 + p["y"][d0:d1] ** 2) + 1e-3 * KPC
    p["vx"][d0:d1] += -vc[d0:d1] * p["y"][d0:d1] / Rxy
    p["vy"][d0:d1] += vc[d0:d1] * p["x"][d0:d1] / Rxy
    p["U"][p["type"] == 2] = u_gas
    if rotation is not None:
        # the reference's commented-out merger recipe (Simulation.cpp:57-90): rotate about x, y, z
        ax_, ay_, az_ = rotation
        for a_, b_, c_ in (("x", "y", "z"), ("vx", "vy", "vz")):
            X, Y, Z = p[a_], p[b_], p[c_]
            y1 = Y * np.cos(ax_) - Z * np.sin(ax_); z1 = Y * np.sin(ax_) + Z * np.cos(ax_)
            x2 = X * np.cos(ay_) + z1 * np.sin(ay_); z2 = -X * np.sin(ay_) + z1 * np.cos(ay_)
            x3 = x2 * np.cos(az_) - y1 * np.sin(az_); y3 = x2 * np.sin(az_) + y1 * np.cos(az_)
            p[a_], p[b_], p[c_] = x3, y3, z2
    for k, c in zip(("x", "y", "z"), centre):
        p[k] = p[k] + c
    for k, c in zip(("vx", "vy", "vz"), bulk_velocity):
        p[k] = p[k] + c
    return p


def merger(n, seed=1234, gas_disk_fraction=0.5):
    """C4: two gas-rich disks; the second rotated by (0.9, 2.2, -1.14) rad, offset (0.2, 0.1, 0) Mpc,
    velocity (-1000, -500, 0) km/s (Simulation.cpp:57-90)."""
    a = disk_galaxy(n // 2, seed, gas_disk_fraction)
    b = disk_galaxy(n - n // 2, seed + 1, gas_disk_fraction, centre=(0.2e3 * KPC, 0.1e3 * KPC, 0.0),
                    bulk_velocity=(-1.0e6, -0.5e6, 0.0), rotation=(0.9, 2.2, -1.14))
    return {k: np.concatenate([a[k], b[k]]) for k in a}


def gas_mass_in_h(p, neighbours=64):
    """massInH = `neighbours` gas-particle masses (64 m_gas for C2/C3)."""
    g = p["type"] == 2
    return float(neighbours * p["mass"][g].mean()) if g.any() else 1e40
