"""oracle/extended.py — numpy / scipy restatement of the extended-accuracy mode (agb_extended.cu).  TEST INFRASTRUCTURE ONLY.

The reference has no such mode (parity unpinned by the reference); these are the textbook formulas the
CUDA kernels implement, evaluated by brute force: direct summation for gravity (cubic-spline softened Newtonian, Springel,
Yoshida & White 2001, softening length h_s = 2.8 eps) and cKDTree neighbour lists for SPH (smoothing length per particle
from (4 pi / 3) (2 h)^3 rho(h) = massInH, rho = sum_j m_j W(r_ij, h_i); "gather" pressure + Monaghan-Gingold viscosity)."""
import numpy as np
from scipy.optimize import brentq
from scipy.spatial import cKDTree

G = 6.67430e-11
GAMMA, KB, PRTN = 5.0 / 3.0, 1.38064852e-23, 1.6726219e-27


def soft_fac(r2, eps):
    hs = 2.8 * eps
    r = np.sqrt(r2)
    u = r / hs
    with np.errstate(divide="ignore", invalid="ignore"):
        out = np.where(r >= hs, 1.0 / (r2 * r), 0.0)
        inner = (10.666666666666666 + u * u * (32.0 * u - 38.4)) / hs**3
        mid = (21.333333333333332 - 48.0 * u + 38.4 * u * u - 10.666666666666666 * u**3 - 0.06666666666666667 / u**3) / hs**3
    out = np.where(u < 0.5, inner, np.where(u < 1.0, mid, out))
    return np.where(r2 == 0.0, 0.0, out)


def direct_gravity(p, targets, eps):
    """acceleration of `targets` (indices) from every particle, spline-softened Newtonian"""
    x = np.stack([p["x"], p["y"], p["z"]], 1)
    m = p["mass"]
    acc = np.zeros((len(targets), 3))
    for k, t in enumerate(targets):
        d = x[t] - x
        r2 = (d * d).sum(1)
        f = m * soft_fac(r2, eps)
        acc[k] = -G * (f[:, None] * d).sum(0)
    return acc


def W(r, h):
    a = 1.0 / (np.pi * h**3)
    q = r / h
    return a * np.where(q < 1.0, 1 - 1.5 * q * q + 0.75 * q**3, np.where(q < 2.0, 0.25 * (2 - q) ** 3, 0.0))


def dW(r, h):
    a = 1.0 / (np.pi * h**4)
    q = r / h
    return a * np.where(q < 1.0, -3.0 * q + 2.25 * q * q, np.where(q < 2.0, -0.75 * (2 - q) ** 2, 0.0))


def sph_density(p, massInH):
    """h, rho, P, T of every gas particle (0 elsewhere)"""
    gas = np.flatnonzero(p["type"] == 2)
    x = np.stack([p["x"], p["y"], p["z"]], 1)[gas]
    m = p["mass"][gas]
    tree = cKDTree(x)
    n = len(p["x"])
    h = np.zeros(n); rho = np.zeros(n)
    span = float(np.abs(x).max()) * 4

    def rho_at(i, hh):
        idx = tree.query_ball_point(x[i], 2 * hh)
        r = np.linalg.norm(x[idx] - x[i], axis=1)
        return float((m[idx] * W(r, hh)).sum())
    for i in range(len(gas)):
        F = lambda hh: (4 * np.pi / 3) * 8 * hh**3 * rho_at(i, hh) - massInH
        lo = hi = span * 1e-3
        while F(lo) > 0:
            lo *= 0.5
        while F(hi) <= 0 and hi < 8 * span:
            hi *= 2
        hh = brentq(F, lo, hi, xtol=lo * 1e-13, rtol=1e-13)
        h[gas[i]] = hh
        rho[gas[i]] = rho_at(i, hh)
    P = (GAMMA - 1.0) * p["U"] * rho
    T = np.where(p["type"] == 2, (GAMMA - 1.0) * p["U"] * PRTN * p["mu"] / KB, 0.0)
    return h, rho, P, T


def sph_forces(p, h, rho, P):
    """(acc [n,3], dUdt [n]) of the gas particles: gather form with the target's kernel, alpha = 0.5, beta = 1, eta^2 = 0.01"""
    gas = np.flatnonzero(p["type"] == 2)
    x = np.stack([p["x"], p["y"], p["z"]], 1)
    v = np.stack([p["vx"], p["vy"], p["vz"]], 1)
    tree = cKDTree(x[gas])
    n = len(p["x"])
    acc = np.zeros((n, 3)); dU = np.zeros(n)
    for i in gas:
        if not (h[i] > 0 and rho[i] > 0):
            continue
        nb = gas[tree.query_ball_point(x[i], 2 * h[i])]
        nb = nb[nb != i]
        d = x[i] - x[nb]
        r2 = (d * d).sum(1)
        keep = (r2 < 4 * h[i] ** 2) & (r2 > 0) & (rho[nb] > 0)
        nb, d, r2 = nb[keep], d[keep], r2[keep]
        r = np.sqrt(r2)
        vij = v[i] - v[nb]
        vr = (vij * d).sum(1)
        ci = np.sqrt(GAMMA * P[i] / rho[i]); cj = np.sqrt(GAMMA * P[nb] / rho[nb])
        hij = 0.5 * (h[i] + h[nb]); cij = 0.5 * (ci + cj); rij = 0.5 * (rho[i] + rho[nb])
        mu = hij * vr / (r2 + 0.01 * hij**2)
        visc = np.where(vr < 0, (-0.5 * cij * mu + mu * mu) / rij, 0.0)
        term = p["mass"][nb] * (P[i] / rho[i] ** 2 + P[nb] / rho[nb] ** 2 + visc) * dW(r, h[i]) / r
        acc[i] = -(term[:, None] * d).sum(0)
        dU[i] = 0.5 * (term * vr).sum()
    return acc, dU
