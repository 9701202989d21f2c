"""Offline study (CPU, numpy) of how the choice of 32-target groups changes the work of the warp-cooperative walk:
list entries per target, lane fill of the entries, straddling pops.  The walk's decisions are the reference's per-target
MAC (radius / r < theta, radius = cell half-width, r = distance to the cell's centre of mass; leaves always accepted), so
the accepted sets are the same for every grouping -- only their packaging into (source, lane-mask) entries changes.

    python tools/group_study.py [n] [theta]

Groupings compared (targets always in tree order unless stated):
  fixed32      32 consecutive targets (what k_walk does)
  cut28..32    boundaries moved to the shallowest cell boundary inside a window, sizes 24..32
  nodes<=32    maximal tree cells with <= 32 particles, greedily merged with following siblings while the sum stays <= 32
  hilbert32    32 consecutive targets along a Hilbert curve (same tree)
  fixed16/64   other group sizes (64 = two targets per lane: an entry then costs ~1.7x, a straddling test 2x)
"""
import sys
import numpy as np

sys.setrecursionlimit(10000)


def plummer(n, seed=1234, a=1.0):
    rng = np.random.default_rng(seed)
    x = rng.random(n)
    r = a / np.sqrt(x ** (-2.0 / 3.0) - 1.0)
    u = rng.normal(size=(n, 3)); u /= np.linalg.norm(u, axis=1)[:, None]
    return r[:, None] * u


def build_tree(pos, depth=21):
    """Octree over the cube [-R, R]^3, one particle per leaf, every level kept.  Returns sorted positions and node arrays."""
    d = np.linalg.norm(pos, axis=1)
    lim = d.mean() + 10 * d.std()
    R = d[d <= lim].max()
    keep = np.all(np.abs(pos) <= R, axis=1)
    pos = pos[keep]
    q = np.minimum(((pos + R) / (2 * R) * (1 << depth)).astype(np.int64), (1 << depth) - 1)
    key = np.zeros(len(pos), np.int64)
    for l in range(depth):
        sh = depth - 1 - l
        key = (key << 3) | (((q[:, 0] >> sh) & 1) | (((q[:, 1] >> sh) & 1) << 1) | (((q[:, 2] >> sh) & 1) << 2))
    order = np.argsort(key, kind="stable")
    pos, key = pos[order], key[order]
    nodes = {"level": [], "first": [], "last": [], "children": []}

    def rec(first, last, level):
        k = len(nodes["level"])
        nodes["level"].append(level); nodes["first"].append(first); nodes["last"].append(last); nodes["children"].append([])
        sh = 3 * (depth - 1 - level)
        dig = (key[first:last + 1] >> sh) & 7
        b = np.searchsorted(dig, np.arange(9))
        for o in range(8):
            a0, a1 = first + b[o], first + b[o + 1] - 1
            if a1 < a0:
                continue
            if a1 == a0:
                nodes["children"][k].append(-(a0 + 1))          # leaf: particle a0
            else:
                nodes["children"][k].append(rec(a0, a1, level + 1))
        return k
    rec(0, len(pos) - 1, 0)
    M = len(nodes["level"])
    level = np.array(nodes["level"]); first = np.array(nodes["first"]); last = np.array(nodes["last"])
    csum = np.vstack([np.zeros(3), np.cumsum(pos, axis=0)])
    com = (csum[last + 1] - csum[first]) / (last - first + 1)[:, None]          # equal masses
    radius = R / (2.0 ** level)
    return pos, key, R, dict(level=level, first=first, last=last, com=com, radius=radius, children=nodes["children"], M=M)


def walk_group(tpos, T, theta):
    """Entries, pairs, pops, straddling pops of one group of targets (tpos: [k, 3]) against tree T."""
    k = len(tpos)
    stack = [(0, np.ones(k, bool))]
    entries = pairs = pops = straddle = 0
    com, radius, children = T["com"], T["radius"], T["children"]
    while stack:
        node, sel = stack.pop()
        pops += 1
        r = np.linalg.norm(com[node] - tpos, axis=1)
        acc = sel & (radius[node] < theta * r)
        opn = sel & ~acc
        na, no = int(acc.sum()), int(opn.sum())
        if na and no:
            straddle += 1
        if na:
            entries += 1; pairs += na
        if no:
            for ch in children[node]:
                if ch < 0:
                    entries += 1; pairs += no                  # leaf (the own leaf is masked out in the kernel; negligible here)
                else:
                    stack.append((ch, opn))
    return entries, pairs, pops, straddle


def hilbert_order(pos, R, bits=16):
    """Skilling's transpose-to-Hilbert on integer coordinates; returns the permutation that sorts along the curve."""
    X = np.minimum(((pos + R) / (2 * R) * (1 << bits)).astype(np.int64), (1 << bits) - 1).T.copy()
    n = 3
    M = 1 << (bits - 1)
    Q = M
    while Q > 1:
        P = Q - 1
        for i in range(n):
            hi = (X[i] & Q) != 0
            X[0] = np.where(hi, X[0] ^ P, X[0])
            t = np.where(~hi, (X[0] ^ X[i]) & P, 0)
            X[0] ^= t; X[i] ^= t
        Q >>= 1
    for i in range(1, n):
        X[i] ^= X[i - 1]
    t = np.zeros_like(X[0])
    Q = M
    while Q > 1:
        t = np.where((X[n - 1] & Q) != 0, t ^ (Q - 1), t)
        Q >>= 1
    for i in range(n):
        X[i] ^= t
    h = np.zeros(pos.shape[0], dtype=object)
    for b in range(bits - 1, -1, -1):
        for i in range(n):
            h = h * 2 + ((X[i] >> b) & 1)
    return np.argsort(h, kind="stable")


def groups_fixed(n, size=32):
    return [np.arange(i, min(n, i + size)) for i in range(0, n, size)]


def groups_cut(key, depth, lo=24, hi=32):
    """Greedy: each group takes between lo and hi targets, ending at the shallowest cell boundary available in that window."""
    n = len(key)
    x = key[1:] ^ key[:-1]
    lcp = np.array([depth - (int(v).bit_length() + 2) // 3 for v in x])      # common levels of neighbours i, i+1
    out, i = [], 0
    while i < n:
        if n - i <= hi:
            out.append(np.arange(i, n)); break
        cand = np.arange(i + lo - 1, i + hi)                                    # last index of the group
        j = cand[np.argmin(lcp[cand])]
        out.append(np.arange(i, j + 1)); i = j + 1
    return out


def groups_nodes(T, n, cap=32):
    first, last, children = T["first"], T["last"], T["children"]
    cells = []

    def rec(node):
        if last[node] - first[node] + 1 <= cap:
            cells.append((first[node], last[node])); return
        for ch in children[node]:
            if ch < 0:
                cells.append((-ch - 1, -ch - 1))
            else:
                rec(ch)
    rec(0)
    out, cur = [], None
    for a, b in cells:                                                          # merge neighbours while they fit
        if cur is not None and b - cur[0] + 1 <= cap:
            cur = (cur[0], b)
        else:
            if cur is not None:
                out.append(np.arange(cur[0], cur[1] + 1))
            cur = (a, b)
    out.append(np.arange(cur[0], cur[1] + 1))
    return out


def evaluate(name, groups, pos, T, theta, sample, rng):
    pick = rng.choice(len(groups), size=min(sample, len(groups)), replace=False)
    tot = np.zeros(4); ntar = 0
    for g in pick:
        idx = groups[g]
        tot += walk_group(pos[idx], T, theta); ntar += len(idx)
    e, p, pops, st = tot
    ng = len(pick)
    fill = ntar / (32.0 * ng)
    # cost model from the ncu profile of k_walk on C1: ~20 warp instructions per list entry (pair loops, per 32 lanes),
    # ~25 per pop, ~25 per straddling pop; per TARGET (a group always costs a full warp)
    per_lane = max(1.0, np.ceil(max(len(groups[g]) for g in pick) / 32.0))      # targets per lane (64-target groups: 2)
    cost = (20 * (0.3 + 0.7 * per_lane) * e + 25 * pops + 25 * per_lane * st) / ntar
    print("%-10s groups %6d  targets/group %5.1f  entries/group %6.1f  pairs/entry %5.2f  pops/group %6.1f  straddling %4.1f%%  interactions/target %6.1f  model cost/target %6.1f"
          % (name, len(groups), 32 * fill, e / ng, p / e, pops / ng, 100 * st / pops, p / ntar, cost))
    return cost


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
    theta = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
    depth = 21
    pos, key, R, T = build_tree(plummer(n), depth)
    n = len(pos)
    print("particles in tree", n, "nodes", T["M"], "theta", theta)
    rng = np.random.default_rng(1)
    sample = 400
    base = evaluate("fixed32", groups_fixed(n), pos, T, theta, sample, np.random.default_rng(1))
    for lo in (28,):
        evaluate("cut%d..32" % lo, groups_cut(key, depth, lo, 32), pos, T, theta, sample, np.random.default_rng(1))
    evaluate("nodes<=32", groups_nodes(T, n), pos, T, theta, sample, np.random.default_rng(1))
    h = hilbert_order(pos, R)
    hg = [h[i:i + 32] for i in range(0, n, 32)]
    evaluate("hilbert32", hg, pos, T, theta, sample, np.random.default_rng(1))
    rank = np.empty(n, np.int64); rank[h] = np.arange(n)                        # position along the Hilbert curve
    for blk in (256, 2048):                                                     # Hilbert order only inside blocks of tree-ordered targets
        lg = []
        for b0 in range(0, n, blk):
            idx = np.arange(b0, min(n, b0 + blk))
            idx = idx[np.argsort(rank[idx], kind="stable")]
            lg += [idx[i:i + 32] for i in range(0, len(idx), 32)]
        evaluate("hilb/%d" % blk, lg, pos, T, theta, sample, np.random.default_rng(1))
    evaluate("fixed16", groups_fixed(n, 16), pos, T, theta, sample, np.random.default_rng(1))
    evaluate("fixed64", groups_fixed(n, 64), pos, T, theta, sample // 2, np.random.default_rng(1))       # two targets per lane
    evaluate("hilbert64", [h[i:i + 64] for i in range(0, n, 64)], pos, T, theta, sample // 2, np.random.default_rng(1))


if __name__ == "__main__":
    main()
