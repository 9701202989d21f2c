"""Multi-GPU plumbing of the force path: one process per GPU, every rank owns a
contiguous 1/G of the particles, the particle arrays are all-gathered once per step (the path's only
collective), every GPU builds the same tree and walks one tree-ordered slice of the targets.

Pure torch.distributed, so the same code runs on NCCL (GPUs) and on gloo (CPU tests)."""
import torch
import torch.distributed as dist

GROUP = 256         # targets per far-field super-group (8 warps); slices are multiples of it => bit-identical results for any world size


def shard_bounds(n, rank, world):
    """Caller-order ownership: particles [lo, hi) belong to `rank`."""
    return n * rank // world, n * (rank + 1) // world


def shard_counts(n, world):
    return [shard_bounds(n, r, world)[1] - shard_bounds(n, r, world)[0] for r in range(world)]


def slice_bounds(n, part, nparts):
    """Tree-order target slice walked by `part` — the rule agb_forces_slice applies (agb_api.cu)."""
    ngrp = (n + GROUP - 1) // GROUP
    return min(n, ngrp * part // nparts * GROUP), min(n, ngrp * (part + 1) // nparts * GROUP)


def gather_particles(shard, n, world, out=None, scratch=None):
    """All-gather a dict of rank-local 1-D tensors into full-length tensors.  Shards may differ by one
    element; they are padded to a common length so that one all_gather_into_tensor per array suffices on
    both NCCL and gloo, then compacted."""
    counts = shard_counts(n, world)
    full = out if out is not None else {k: torch.empty(n, dtype=v.dtype, device=v.device) for k, v in shard.items()}
    if world == 1:
        for k, v in shard.items():
            full[k].copy_(v)
        return full
    mx = max(counts)
    even = all(c == mx for c in counts)
    for k, v in shard.items():
        if even:
            dist.all_gather_into_tensor(full[k], v.contiguous())
            continue
        pad = torch.zeros(mx, dtype=v.dtype, device=v.device) if scratch is None else scratch.setdefault((k, "pad"), torch.zeros(mx, dtype=v.dtype, device=v.device))
        pad[: v.numel()].copy_(v)
        buf = torch.empty(world * mx, dtype=v.dtype, device=v.device) if scratch is None else scratch.setdefault((k, "buf"), torch.empty(world * mx, dtype=v.dtype, device=v.device))
        dist.all_gather_into_tensor(buf, pad)
        off = 0
        for r, c in enumerate(counts):
            full[k][off: off + c].copy_(buf[r * mx: r * mx + c])
            off += c
    return full
