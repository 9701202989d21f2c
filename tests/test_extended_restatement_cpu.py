"""CPU checks of oracle/extended.py, the numpy restatement the extended-accuracy GPU tests compare with (test infrastructure):
closed-form cases of the spline-softened attraction and of the SPH kernel sums."""
import numpy as np

from oracle import extended


def test_spline_softening_limits():
    eps = 1.0
    hs = 2.8 * eps
    r = np.array([1e-3, 0.5 * hs, 0.999999 * hs, hs, 3.0 * hs])
    f = extended.soft_fac(r * r, eps)
    assert np.allclose(f[3:], 1.0 / r[3:] ** 3, rtol=1e-14)                   # Newtonian outside the softening length
    assert np.isclose(f[2], 1.0 / hs**3, rtol=1e-4)                           # continuous at r = h_s
    assert np.isclose(f[0], 10.666666666666666 / hs**3, rtol=1e-5)           # harmonic core: a = -(32/3) G m d / h_s^3
    lo, hi = extended.soft_fac(np.array([(0.5 * hs * (1 - 1e-9)) ** 2, (0.5 * hs * (1 + 1e-9)) ** 2]), eps)
    assert np.isclose(lo, hi, rtol=1e-7)                                      # continuous at r = h_s / 2
    assert extended.soft_fac(np.array([0.0]), eps)[0] == 0.0                  # the particle itself


def test_two_body_direct_sum():
    p = {"x": np.array([0.0, 10.0]), "y": np.zeros(2), "z": np.zeros(2), "mass": np.array([3.0, 5.0])}
    a = extended.direct_gravity(p, [0, 1], eps=0.1)
    assert np.allclose(a[0], [extended.G * 5.0 / 100.0, 0, 0]) and np.allclose(a[1], [-extended.G * 3.0 / 100.0, 0, 0])


def test_kernel_normalisation_and_lattice_density():
    # W integrates to one: sum over a fine lattice
    h = 1.0
    g = np.arange(-2.05, 2.1, 0.1)
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    r = np.sqrt(X * X + Y * Y + Z * Z).ravel()
    assert abs(extended.W(r, h).sum() * 0.1**3 - 1.0) < 2e-3
    # a uniform lattice of gas: the solved smoothing lengths enclose massInH and the density is the lattice's
    n = 9
    c = (np.arange(n) - n // 2).astype(float)
    X, Y, Z = np.meshgrid(c, c, c, indexing="ij")
    N = n**3
    p = {"x": X.ravel(), "y": Y.ravel(), "z": Z.ravel(), "mass": np.ones(N), "type": np.full(N, 2, np.uint8), "U": np.ones(N), "mu": np.full(N, 0.58),
         "vx": np.zeros(N), "vy": np.zeros(N), "vz": np.zeros(N)}
    h, rho, P, T = extended.sph_density(p, 40.0)
    mid = N // 2
    assert abs((4 * np.pi / 3) * 8 * h[mid] ** 3 * rho[mid] - 40.0) < 1e-9
    assert abs(rho[mid] - 1.0) < 0.05                                         # one unit mass per unit cell
    acc, dU = extended.sph_forces(p, h, rho, P)
    assert np.linalg.norm(acc[mid]) < 1e-12 * abs(P[mid])                      # symmetric neighbourhood: no net pressure force, no heating at rest
    assert dU[mid] == 0.0
