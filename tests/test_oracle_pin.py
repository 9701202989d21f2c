"""CPU tests (-m "not gpu"): the plain-C oracle restatement is pinned bit-for-bit (i) against golden
vectors produced by the unmodified reference on its own shipped example ICs and (ii), where the
reference has been compiled in place (oracle/_ref), against the reference itself on synthetic sets."""
import numpy as np
import pytest

from conftest import GOLDEN_CASES, load_golden
from parity import adversarial_set

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
    p, args = adversarial_set(pkg, seed)
    n, cores = len(p["x"]), args[4]
    got = oracle.run(p, *args)
    want = oracle.run_ref(p, *args)
    assert got["R"] == want["R"]
    for k in BITWISE:
        assert np.array_equal(got[k], want[k], equal_nan=True), (k, n, cores)
    for k, v in want["nodes"].items():
        assert np.array_equal(got["nodes"][k], v, equal_nan=True), ("node." + k, n, cores)
