"""World-size-2 gloo tests (CPU) of the N > 1 host logic: particle sharding, the per-step all-gather and
the tree-order slice rule that agb_forces_slice applies."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    pkg = ge.load_package()
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        p = pkg.ics.plummer(n, seed=5, gas_fraction=0.3)
        lo, hi = pkg.shard.shard_bounds(n, rank, world)
        shard = {k: torch.from_numpy(np.ascontiguousarray(p[k][lo:hi])) for k in ("x", "y", "z", "mass", "vx", "type")}
        full = pkg.shard.gather_particles(shard, n, world)
        ok = all(np.array_equal(full[k].numpy(), p[k]) for k in shard)
        pg = pkg.shard.PackedGather(shard, n, world)
        out = pg.gather(shard)
        ok = ok and all(np.array_equal(out[k].numpy(), p[k]) for k in shard)
        if n % world == 0:
            ig = pkg.shard.InPlaceGather({k: v.dtype for k, v in shard.items()}, n, world, rank, torch.device("cpu"))
            out2 = ig.gather(shard)                    # copies the shard into its slot, then completes the arrays in place
            ok = ok and all(np.array_equal(out2[k].numpy(), p[k]) for k in shard)
            ok = ok and all(ig.shard[k].data_ptr() == out2[k][lo:hi].data_ptr() for k in shard)
            # the staged exchange of the bench: one group of arrays at a time; a group completes exactly its own arrays
            ig2 = pkg.shard.InPlaceGather({k: v.dtype for k, v in shard.items()}, n, world, rank, torch.device("cpu"))
            for k, v in ig2.out.items():
                v.fill_(0)
            for k in shard:
                ig2.shard[k].copy_(shard[k])
            first, rest = ("x", "y", "z", "mass", "type"), ("vx",)
            ig2.gather_fields(first)
            ok = ok and all(np.array_equal(ig2.out[k].numpy(), p[k]) for k in first)
            other = np.zeros(n, p["vx"].dtype); other[lo:hi] = p["vx"][lo:hi]
            ok = ok and np.array_equal(ig2.out["vx"].numpy(), other)        # not exchanged yet: only this rank's slot is filled
            ig2.gather_fields(rest)
            ok = ok and np.array_equal(ig2.out["vx"].numpy(), p["vx"])
        q.put((rank, ok, pkg.shard.slice_bounds(n, rank, world)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [1000, 4099])
def test_gather_and_slices_world2(n):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res)
    (a0, a1), (b0, b1) = res[0][2], res[1][2]
    assert a0 == 0 and a1 == b0 and b1 == n and a1 % 256 == 0          # contiguous cover, boundaries on warp groups


def test_slice_rule_covers_everything(pkg):
    for n in (0, 1, 31, 32, 33, 1000, 1_000_000, 16_000_001):
        for parts in (1, 2, 4, 8):
            b = [pkg.shard.slice_bounds(n, r, parts) for r in range(parts)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(parts - 1))
            assert all(x[0] % 256 == 0 for x in b)
            assert sum(pkg.shard.shard_counts(n, parts)) == n
