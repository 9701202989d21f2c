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
    """Slice of the (active, tree-ordered) targets walked by `part` — the rule agb_forces_slice applies on the device
    (target_slice in agb_walk.cu); n = number of active targets."""
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


class PackedGather:
    """One NCCL all-gather per step for all float64 particle arrays (plus one for the uint8 types): every rank keeps its
    shard as rows of a packed [fields, max_count] buffer; the gathered [world, fields, max_count] block is transposed on
    the device into full-length per-field arrays.  5 collectives per step become 2 (latency bound at 1M particles)."""

    def __init__(self, shard, n, world):
        self.n, self.world = n, world
        self.counts = shard_counts(n, world)
        self.mx = max(self.counts)
        self.fields = [k for k, v in shard.items() if v.dtype == torch.float64]
        self.bytes_fields = [k for k, v in shard.items() if v.dtype == torch.uint8]
        dev = next(iter(shard.values())).device
        self.local = torch.zeros(len(self.fields), self.mx, dtype=torch.float64, device=dev)
        self.local_b = torch.zeros(max(1, len(self.bytes_fields)), self.mx, dtype=torch.uint8, device=dev)
        self.recv = torch.empty(world, len(self.fields), self.mx, dtype=torch.float64, device=dev)
        self.recv_b = torch.empty(world, max(1, len(self.bytes_fields)), self.mx, dtype=torch.uint8, device=dev)
        self.full = torch.empty(len(self.fields), world * self.mx, dtype=torch.float64, device=dev)
        self.full_b = torch.empty(max(1, len(self.bytes_fields)), world * self.mx, dtype=torch.uint8, device=dev)
        self.even = all(c == self.mx for c in self.counts)
        self.out = {k: (self.full[i, :n] if self.even else torch.empty(n, dtype=torch.float64, device=dev)) for i, k in enumerate(self.fields)}
        self.out.update({k: (self.full_b[i, :n] if self.even else torch.empty(n, dtype=torch.uint8, device=dev)) for i, k in enumerate(self.bytes_fields)})

    def gather(self, shard):
        c = next(iter(shard.values())).numel()
        for i, k in enumerate(self.fields):
            self.local[i, :c].copy_(shard[k])
        for i, k in enumerate(self.bytes_fields):
            self.local_b[i, :c].copy_(shard[k])
        dist.all_gather_into_tensor(self.recv.view(-1), self.local.view(-1))
        self.full.view(len(self.fields), self.world, self.mx).copy_(self.recv.permute(1, 0, 2))
        if self.bytes_fields:
            dist.all_gather_into_tensor(self.recv_b.view(-1), self.local_b.view(-1))
            self.full_b.view(len(self.bytes_fields), self.world, self.mx).copy_(self.recv_b.permute(1, 0, 2))
        if not self.even:
            for i, k in enumerate(self.fields + self.bytes_fields):
                src = self.full if i < len(self.fields) else self.full_b
                row = i if i < len(self.fields) else i - len(self.fields)
                off = 0
                for r, cnt in enumerate(self.counts):
                    self.out[k][off: off + cnt].copy_(src[row, r * self.mx: r * self.mx + cnt])
                    off += cnt
        return self.out


class InPlaceGather:
    """The cheapest form of the step's exchange: every rank's shard IS its slot of the full-length arrays (a distributed
    driver integrates it in place), and the arrays are completed by in-place all-gathers issued as ONE coalesced NCCL group —
    no packing, no transposes, one launch.  Needs n % world == 0 (equal slots); PackedGather covers the ragged case."""

    def __init__(self, dtypes, n, world, rank, device):
        if n % world:
            raise ValueError("InPlaceGather needs n divisible by the world size")
        self.n, self.world, self.rank, self.device = n, world, rank, device
        lo, hi = shard_bounds(n, rank, world)
        self.out = {k: torch.empty(n, dtype=dt, device=device) for k, dt in dtypes.items()}
        self.shard = {k: v[lo:hi] for k, v in self.out.items()}

    def gather_fields(self, fields):
        """One coalesced group of in-place all-gathers for some of the arrays (a driver that hands the arrays over in stages:
        positions / masses / types first, so that the tree build overlaps the exchange of the rest)."""
        if self.world > 1:
            coalesce = getattr(dist, "_coalescing_manager", None)
            if coalesce is not None and dist.get_backend() == "nccl":
                with coalesce(device=self.device):
                    for k in fields:
                        dist.all_gather_into_tensor(self.out[k], self.shard[k])
            else:
                for k in fields:
                    dist.all_gather_into_tensor(self.out[k], self.shard[k].clone())
        return self.out

    def gather(self, shard=None):
        if shard is not None and shard is not self.shard:
            for k, v in shard.items():
                if v.data_ptr() != self.shard[k].data_ptr():
                    self.shard[k].copy_(v)
        if self.world > 1:
            coalesce = getattr(dist, "_coalescing_manager", None)
            if coalesce is not None and dist.get_backend() == "nccl":
                with coalesce(device=self.device):
                    for k in self.out:
                        dist.all_gather_into_tensor(self.out[k], self.shard[k])
            else:
                for k in self.out:
                    dist.all_gather_into_tensor(self.out[k], self.shard[k].clone())
        return self.out
