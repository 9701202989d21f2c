"""Parity tiers of SURVEY.md §8(c): compares the GPU path (through the C ABI) with an oracle result
(oracle/oracle.py run() / run_ref() / a golden fixture).  Returns a report dict; assert_parity raises."""
import numpy as np

ACC_MEDIAN_TOL = 1e-6      # north_star: median relative acceleration error
ACC_P99_TOL = 1e-4         # north_star: 99th percentile
NODE_RTOL = 1e-13          # node mass / COM / gasMass / mVel (summation order differs)
RHO_RTOL = 1e-12           # FP64 segmented sum vs the reference's serial sum


def relerr_vec(a, b):
    num = np.sqrt(sum((x - y) ** 2 for x, y in zip(a, b)))
    den = np.sqrt(sum(y ** 2 for y in b))
    ok = den > 0
    out = np.zeros_like(num)
    out[ok] = num[ok] / den[ok]
    out[~ok] = np.where(num[~ok] == 0, 0.0, np.inf)
    return out


def compare(got, want, p, ctx=None, check_counters=True, check_nodes=True):
    rep = {}
    gas = np.asarray(p["type"]) == 2
    rep["R_equal"] = got["R"] == want["R"]
    rel = relerr_vec((got["ax"], got["ay"], got["az"]), (want["ax"], want["ay"], want["az"]))
    rep["acc_median"] = float(np.median(rel)); rep["acc_p99"] = float(np.percentile(rel, 99)); rep["acc_max"] = float(rel.max())
    rep["h_mismatch"] = int((got["h"][gas] != want["h"][gas]).sum())
    rep["vis_maxrel"] = float(np.max(np.abs(got["vis"] - want["vis"]) / np.where(want["vis"] != 0, np.abs(want["vis"]), 1.0))) if len(rel) else 0.0
    rep["vis_zero_mismatch"] = int(((got["vis"] == 0) != (want["vis"] == 0)).sum())
    for k in ("rho", "P", "T"):
        w = want[k][gas]; g = got[k][gas]
        rep[k + "_maxrel"] = float(np.max(np.abs(g - w) / np.where(w != 0, np.abs(w), 1.0))) if gas.any() else 0.0
    w = want["dUdt"]; g = got["dUdt"]
    nz = w != 0
    du = np.abs(g - w)[nz] / np.abs(w[nz]) if nz.any() else np.zeros(1)
    rep["dUdt_median"] = float(np.median(du)); rep["dUdt_p99"] = float(np.percentile(du, 99)); rep["dUdt_zero_mismatch"] = int(((g != 0) != nz).sum())
    if ctx is not None:
        ld, hi, lo = ctx.tree_particles()
        rep["leafdepth_mismatch"] = int((ld != want["leafdepth"]).sum())
        rep["key_mismatch"] = int(((hi != want["key_hi"]) | (lo != want["key_lo"])).sum())
        if check_counters:
            tc = ctx.target_counters()
            for k in ("visits", "acc_nodes", "acc_leaves", "sph"):
                rep[k + "_mismatch"] = int((tc[k] != want[k]).sum())
            c = ctx.counters()
            rep["interactions_total_equal"] = c["interactions"] == int(want["acc_nodes"].sum() + want["acc_leaves"].sum())
            rep["exact_fallbacks"] = c["mac_exact_fallbacks"]
        if check_nodes and "nodes" in want:
            nd = ctx.nodes(); wn = want["nodes"]
            internal = wn["isLeaf"] == 0
            wkey = np.stack([wn["depth"][internal].astype(np.uint64), wn["key_hi"][internal], wn["key_lo"][internal]], 1)
            gkey = np.stack([nd["depth"].astype(np.uint64), nd["key_hi"], nd["key_lo"]], 1)
            wo = np.lexsort(wkey.T[::-1]); go = np.lexsort(gkey.T[::-1])
            rep["node_count_equal"] = len(wo) == len(go)
            if rep["node_count_equal"]:
                rep["node_topology_mismatch"] = int((wkey[wo] != gkey[go]).any(1).sum())
                R = want["R"]
                vmax = max(1e-300, float(np.max(np.abs(np.concatenate([p["vx"], p["vy"], p["vz"]])))))
                def nerr(gk, wk, scale=None):
                    a = nd[gk][go]; b = wn[wk][internal][wo]
                    den = np.abs(b) if scale is None else scale
                    den = np.where(den > 0, den, 1.0)
                    return float(np.max(np.abs(a - b) / den)) if len(a) else 0.0
                # the reference sums a node's particles sequentially: its own rounding error grows like count * 2^-53
                ntol = np.maximum(1.0, nd["count"][go].astype(float)) * 2.0 ** -52
                rep["node_mass_maxrel"] = nerr("mass", "mass") / 1.0
                rep["node_mass_over_tol"] = float(np.max(np.abs(nd["mass"][go] - wn["mass"][internal][wo]) / np.maximum(wn["mass"][internal][wo], 1e-300) / ntol))
                rep["node_gas_over_tol"] = float(np.max(np.abs(nd["gasMass"][go] - wn["gasMass"][internal][wo]) / np.maximum(wn["gasMass"][internal][wo], 1e-300) / ntol))
                rep["node_gas_maxrel"] = nerr("gasMass", "gasMass")
                rep["node_com_maxrel_R"] = max(nerr("com" + c, "com" + c, R) for c in "xyz")
                rep["node_mvel_maxrel_v"] = max(nerr("mv" + c, "mv" + c, vmax) for c in "xyz")
                # childParticles sizes: count, doubled at the bulk->single hand-over nodes; the root keeps none in the bulk path
                cnt = nd["count"][go] * (1 + nd["dup"][go]); wc = wn["nchild"][internal][wo]
                notroot = nd["depth"][go] > 0
                rep["node_nchild_mismatch"] = int((cnt[notroot] != wc[notroot]).sum())
    return rep


def assert_parity(rep):
    assert rep["R_equal"], rep
    assert rep["acc_median"] <= ACC_MEDIAN_TOL and rep["acc_p99"] <= ACC_P99_TOL, rep
    assert rep["h_mismatch"] == 0, rep
    assert rep["vis_zero_mismatch"] == 0 and rep["vis_maxrel"] <= NODE_RTOL * 10, rep
    assert rep["rho_maxrel"] <= RHO_RTOL and rep["P_maxrel"] <= RHO_RTOL and rep["T_maxrel"] <= RHO_RTOL, rep
    assert rep["dUdt_zero_mismatch"] == 0 and rep["dUdt_median"] <= ACC_MEDIAN_TOL and rep["dUdt_p99"] <= ACC_P99_TOL, rep
    for k in ("leafdepth_mismatch", "key_mismatch", "visits_mismatch", "acc_nodes_mismatch", "acc_leaves_mismatch", "sph_mismatch",
              "node_topology_mismatch", "node_nchild_mismatch"):
        if k in rep:
            assert rep[k] == 0, (k, rep)
    for k in ("node_count_equal", "interactions_total_equal"):
        if k in rep:
            assert rep[k], (k, rep)
    for k in ("node_com_maxrel_R", "node_mvel_maxrel_v"):
        if k in rep:
            assert rep[k] <= NODE_RTOL * 10, (k, rep)
    for k in ("node_mass_over_tol", "node_gas_over_tol"):          # <= count * 2^-52 relative (the reference's own sequential-sum error)
        if k in rep:
            assert rep[k] <= 1.0, (k, rep)
    for k in ("node_mass_maxrel", "node_gas_maxrel"):
        if k in rep:
            assert rep[k] <= 1e-10, (k, rep)


def beyond_fp32_law(p, want):
    """True when a set lies outside what the mixed-precision pair law resolves and the library's own guard (root cube vs e0, depth
    > 40: FP64 pair arithmetic) does not catch: targets further than ~40 R from the sources (r^6 in units of (R/2^16)^6 leaves the
    FP32 range, the pair force underflows to zero: only particles far beyond mean + 10 sigma), or pairs 2^-38 R apart inside a
    sparse set whose 32-target groups span the whole cube (float-float positions resolve 2^-47 R).  Measured on the adversarial
    sets (tools/gpu_adversarial.py): every discrete decision stays exact, acc p99 reaches 1e-3 on such a set; DESIGN.md §2."""
    R = float(want["R"])
    far = max(float(np.abs(p[c]).max()) for c in ("x", "y", "z")) > 40.0 * R
    return bool(far or int(want["leafdepth"].max()) >= 40)


def adversarial_set(pkg, seed, lattice=True):
    """Small random particle set built to hit the corners of the path: points on power-of-two planes (the split planes of a cube
    whose half-width is a power of two; lattice=False leaves them out — exact opening-test ties on lattices are a documented
    divergence of the GPU path), tight pairs that share ~30 octree levels, a few far outliers, mixed types, unequal
    masses, resting particles, and every `cores` branch of the insertion.  Returns (particles, (theta, e0, massInH, globalTime, cores))."""
    rng = np.random.default_rng(1000 + seed)
    n = int(rng.integers(1, 420))
    L = 1e20
    kind = rng.integers(0, 4, n)
    if not lattice:
        kind = np.where(kind == 1, 0, kind)
    pos = rng.normal(0.0, 1.0, (3, n)) * L
    grid = np.ldexp(rng.integers(-8, 9, (3, n)).astype(np.float64), -3) * 2.0 ** 66           # multiples of 2^63 up to 2^66 ~ 0.7 L
    pos = np.where(kind == 1, grid, pos)
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
    return p, args
