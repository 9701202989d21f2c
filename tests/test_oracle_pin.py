"""CPU tests (-m "not gpu"): the plain-C oracle restatement is pinned bit-for-bit (i) against golden
vectors produced by the unmodified reference on its own shipped example ICs and (ii), where the
reference has been compiled in place (oracle/_ref), against the reference itself on synthetic sets."""
import numpy as np
import pytest

from conftest import GOLDEN_CASES, load_golden

BITWISE = ("ax", "ay", "az", "dUdt", "h", "rho", "P", "T", "vis", "leafdepth", "key_hi", "key_lo", "visits", "acc_nodes", "acc_leaves", "sph")


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_oracle_matches_reference_golden(oracle, name):
    p, want, par = load_golden(name)
    got = oracle.run(p, par["theta"], par["e0"], par["massInH"], par["globalTime"], int(par["cores"]), nodes="nodes" in want)
    assert got["R"] == want["R"]
    for k in BITWISE:
        assert np.array_equal(got[k], want[k]), k
    if "nodes" in want:
        for k, v in want["nodes"].items():
            assert np.array_equal(got["nodes"][k], v), "node." + k


@pytest.mark.parametrize("case", ["plummer_gas", "disk", "merger", "tiny", "serial_root"])
def test_oracle_matches_reference_binary(oracle, pkg, case):
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built here (reference sources absent)")
    ics = pkg.ics
    if case == "plummer_gas":
        p = ics.plummer(6000, seed=3, gas_fraction=0.25); args = (0.5, 1e18, ics.gas_mass_in_h(p, 24), 0.0, 8)
    elif case == "disk":
        p = ics.disk_galaxy(12000, seed=5); args = (0.6, 1e18, ics.gas_mass_in_h(p, 64), 0.0, 4)
    elif case == "merger":
        p = ics.merger(8000, seed=9); args = (0.4, 1e19, ics.gas_mass_in_h(p, 32), 0.0, 1)
    elif case == "tiny":
        p = ics.plummer(3, seed=1, gas_fraction=1.0); args = (0.5, 1e18, 1e40, 0.0, 8)
    else:
        p = ics.plummer(500, seed=11, gas_fraction=0.5); args = (0.5, 1e18, ics.gas_mass_in_h(p, 8), 0.0, 8)   # N < cores*100: one-by-one from the root
    got = oracle.run(p, *args)
    want = oracle.run_ref(p, *args)
    assert got["R"] == want["R"]
    for k in BITWISE:
        assert np.array_equal(got[k], want[k]), k
    for k, v in want["nodes"].items():
        assert np.array_equal(got["nodes"][k], v), "node." + k


def test_inactive_particles_keep_acc(oracle, pkg):
    p = pkg.ics.plummer(2000, seed=2)
    p["next_time"][::2] = 5.0
    o = oracle.run(p, 0.5, 1e18, 1e40, 0.0, 8, nodes=False)
    assert np.all(o["ax"][::2] == 0) and np.all(o["visits"][::2] == 0)
    assert np.all(o["ax"][1::2] != 0)


def test_direct_sum_bound(oracle, pkg):
    """Tree-vs-direct error of the oracle is the reference's own ~1e-2; sanity bound only."""
    p = pkg.ics.plummer(4000, seed=4)
    e0 = 1e18
    o = oracle.run(p, 0.5, e0, 1e40, 0.0, 8, nodes=False)
    idx = np.arange(0, 4000, 40)
    G = 6.67430e-11
    intree = o["leafdepth"] >= 0
    err = []
    for i in idx:
        dx = p["x"] - p["x"][i]; dy = p["y"] - p["y"][i]; dz = p["z"] - p["z"][i]
        r2 = dx * dx + dy * dy + dz * dz
        m = intree & (r2 > 0)
        f = G * p["mass"][m] / (r2[m] + e0 * e0) / np.sqrt(r2[m])
        a = np.array([(f * dx[m]).sum(), (f * dy[m]).sum(), (f * dz[m]).sum()])
        b = np.array([o["ax"][i], o["ay"][i], o["az"][i]])
        err.append(np.linalg.norm(a - b) / np.linalg.norm(a))
    assert np.mean(err) < 5e-2


@pytest.mark.parametrize("seed", range(16))
def test_oracle_matches_reference_binary_on_adversarial_random_sets(oracle, pkg, seed):
    """Small random sets built to hit the corners: points on power-of-two planes (the split planes of a cube whose half-width is a
    power of two), repeated coordinate values, tight pairs, a few far outliers, mixed types, unequal masses, resting particles, and
    every `cores` branch (bulk insertion, hand-over to one-by-one insertion below cores*100, one-by-one from the root)."""
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built here (reference sources absent)")
    rng = np.random.default_rng(1000 + seed)
    n = int(rng.integers(1, 420))
    L = 1e20
    kind = rng.integers(0, 4, n)
    pos = rng.normal(0.0, 1.0, (3, n)) * L
    lattice = np.ldexp(rng.integers(-8, 9, (3, n)).astype(np.float64), -3) * 2.0 ** 66        # multiples of 2^63 up to 2^66 ~ 0.7 L
    pos = np.where(kind == 1, lattice, pos)
    if n > 4:
        src = rng.integers(0, n, n)
        tight = pos[:, src] * (1.0 + 1e-9 * rng.normal(size=(3, n)))                            # pairs that share ~30 levels
        pos = np.where(kind == 2, tight, pos)
        far = rng.random(n) < 0.01
        pos[:, far] *= 1e3                                                                        # beyond mean + 10 sigma or just inside it
    # truly coincident points send the reference (and its restatement) into an unbounded recursion: keep one of each
    _, first = np.unique(pos.T, axis=0, return_index=True)
    dup = np.ones(n, bool); dup[first] = False
    pos[:, dup] = rng.normal(0.0, 1.0, (3, int(dup.sum()))) * L
    p = pkg.ics._empty(n)
    p["x"], p["y"], p["z"] = pos[0].copy(), pos[1].copy(), pos[2].copy()
    p["type"] = rng.choice(np.array([1, 2, 2, 3], np.uint8), n)
    p["mass"] = 1e35 * np.exp(rng.normal(0.0, 0.5, n)) if seed % 2 else np.full(n, 1e35)
    for c in ("vx", "vy", "vz"):
        p[c] = rng.normal(0.0, 1e5, n)
    p["U"] = np.where(p["type"] == 2, 1e9 * (0.5 + rng.random(n)), 0.0)
    if seed % 3 == 0:
        p["next_time"] = np.where(rng.random(n) < 0.3, 7.0, 0.0)
    gas = p["type"] == 2
    mh = float(rng.integers(2, 12)) * (p["mass"][gas].mean() if gas.any() else 1e35)
    cores = int(rng.choice([1, 2, 8]))
    args = (float(rng.choice([0.3, 0.5, 0.8])), 1e18, mh, 0.0, cores)
    got = oracle.run(p, *args)
    want = oracle.run_ref(p, *args)
    assert got["R"] == want["R"]
    for k in BITWISE:
        assert np.array_equal(got[k], want[k], equal_nan=True), (k, n, cores)
    for k, v in want["nodes"].items():
        assert np.array_equal(got["nodes"][k], v, equal_nan=True), ("node." + k, n, cores)
