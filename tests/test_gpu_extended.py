"""GPU tests of the extended-accuracy mode (AGB_OPT_EXTENDED).  The reference has no such algorithm, so
parity is UNPINNED by it: gravity is checked against direct summation of the same softened law, SPH against a numpy /
cKDTree restatement of the same formulas (oracle/extended.py)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ext_ctx(pkg):
    ctx = pkg.Context(0, 8)
    ctx.set_option(pkg.capi.AGB_OPT_EXTENDED, 1)
    return ctx


def test_quadrupole_walk_against_direct_summation(pkg):
    from oracle import extended
    p = pkg.ics.plummer(30000, seed=61)
    eps = 1e18
    rng = np.random.default_rng(1)
    targets = rng.choice(30000, 400, replace=False)
    want = extended.direct_gravity(p, targets, eps)
    ctx = _ext_ctx(pkg)
    try:
        got, _ = pkg.run_step(dict(p), 0.5, eps, 1e40, 0.0, context=ctx)
        ext = np.stack([got["ax"], got["ay"], got["az"]], 1)[targets]
        rel = np.linalg.norm(ext - want, axis=1) / np.linalg.norm(want, axis=1)
        c = ctx.counters()
        print("extended gravity vs direct sum: median %.2e p99 %.2e max %.2e, %.0f interactions/target" % (np.median(rel), np.percentile(rel, 99), rel.max(), c["interactions"] / 30000))
        assert np.median(rel) < 3e-4 and np.percentile(rel, 99) < 3e-3
        # a tighter opening angle converges to the direct sum
        got2, _ = pkg.run_step(dict(p), 0.2, eps, 1e40, 0.0, context=ctx)
        ext2 = np.stack([got2["ax"], got2["ay"], got2["az"]], 1)[targets]
        rel2 = np.linalg.norm(ext2 - want, axis=1) / np.linalg.norm(want, axis=1)
        print("   theta 0.2: median %.2e p99 %.2e" % (np.median(rel2), np.percentile(rel2, 99)))
        assert np.median(rel2) < np.median(rel) / 5 and np.percentile(rel2, 99) < 2e-4
        # the quadrupole term earns its keep: the same walk with monopoles only is several times further off
        ctx.set_option(pkg.capi.AGB_OPT_EXTENDED, 2)
        got1, _ = pkg.run_step(dict(p), 0.5, eps, 1e40, 0.0, context=ctx)
        mono = np.stack([got1["ax"], got1["ay"], got1["az"]], 1)[targets]
        rel1 = np.linalg.norm(mono - want, axis=1) / np.linalg.norm(want, axis=1)
        print("   monopole only, theta 0.5: median %.2e p99 %.2e" % (np.median(rel1), np.percentile(rel1, 99)))
        assert np.median(rel1) > 3 * np.median(rel)
    finally:
        ctx.close()
    # the parity mode's monopole walk (the reference's law, half-width opening test) is an order of magnitude further from ITS direct sum
    # (tests/test_gpu_parity.py::test_direct_sum_bound_gpu: ~1e-2)


def test_neighbour_loop_sph_against_restatement(pkg):
    from oracle import extended
    p = pkg.ics.plummer(6000, seed=62, gas_fraction=0.4)
    rng = np.random.default_rng(2)
    gas = p["type"] == 2
    p["U"] = np.where(gas, 1e9 * (0.5 + rng.random(6000)), 0.0)
    mh = pkg.ics.gas_mass_in_h(p, 48)
    ctx = _ext_ctx(pkg)
    try:
        got, _ = pkg.run_step(dict(p), 0.5, 1e18, mh, 0.0, context=ctx)
        dry = dict(p); dry["type"] = np.where(gas, 1, p["type"]).astype(np.uint8)      # the same masses without gas: gravity alone
        grav, _ = pkg.run_step(dry, 0.5, 1e18, mh, 0.0, context=ctx)
    finally:
        ctx.close()
    # particles outside the root cube (Tree.cpp:85-117) are not in the tree: neither neighbours nor SPH targets, in any mode
    outside = np.maximum(np.maximum(np.abs(p["x"]), np.abs(p["y"])), np.abs(p["z"])) > got["R"]
    q = dict(p); q["type"] = np.where(outside & gas, 1, p["type"]).astype(np.uint8)
    assert not got["h"][outside].any()
    gas = q["type"] == 2
    h, rho, P, T = extended.sph_density(q, mh)
    acc_sph, dU = extended.sph_forces(q, h, rho, P)
    eh = np.abs(got["h"][gas] - h[gas]) / h[gas]; er = np.abs(got["rho"][gas] - rho[gas]) / rho[gas]
    print("extended SPH: h max rel %.2e, rho max rel %.2e" % (eh.max(), er.max()))
    assert eh.max() < 1e-6 and er.max() < 1e-5
    assert np.allclose(got["P"][gas], P[gas], rtol=1e-5, atol=0) and np.allclose(got["T"][gas], T[gas], rtol=1e-12, atol=0)
    assert not got["h"][~gas].any()
    sph = np.stack([got[k] - grav[k] for k in ("ax", "ay", "az")], 1)
    scale = np.linalg.norm(acc_sph[gas], axis=1)
    err = np.linalg.norm(sph[gas] - acc_sph[gas], axis=1) / np.maximum(scale, np.median(scale) * 1e-3)
    print("extended SPH acc vs restatement: median %.2e max %.2e; neighbours per particle ~%.0f" % (np.median(err), err.max(), mh / p["mass"][gas][0]))
    assert np.median(err) < 1e-5 and np.percentile(err, 99) < 1e-3            # the difference of two accelerations loses digits where gravity dominates
    nz = dU != 0
    ed = np.abs(got["dUdt"][nz] - dU[nz]) / np.maximum(np.abs(dU[nz]), 1e-6 * np.abs(dU).max())
    print("   dU/dt: median %.2e max %.2e" % (np.median(ed), ed.max()))
    assert np.median(ed) < 1e-5 and ed.max() < 1e-2
    assert not got["dUdt"][~gas].any()


def test_extended_mode_slices_and_multi_handle(pkg):
    """The extended kernels walk target slices like the parity kernels: several devices give the same bits as one."""
    import torch
    p = pkg.ics.disk_galaxy(40000, seed=63)
    mh = pkg.ics.gas_mass_in_h(p, 48)
    ctx = _ext_ctx(pkg)
    try:
        want, _ = pkg.run_step(dict(p), 0.5, 1e18, mh, 0.0, context=ctx)
    finally:
        ctx.close()
    m = pkg.MultiContext([0, 1] if torch.cuda.device_count() >= 2 else [0, 0], 8)
    try:
        m.set_option(pkg.capi.AGB_OPT_EXTENDED, 1)
        m.set_particles(dict(p))
        R = m.build_tree(); m.visual_density(R / 1e5); m.gas_density(mh); m.forces(0.0, 1e18, 0.5)
        got = m.results()
    finally:
        m.close()
    for k in ("ax", "ay", "az", "dUdt", "h", "rho", "P", "T"):
        assert np.array_equal(got[k], want[k]), k
