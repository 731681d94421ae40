"""Multi-GPU host logic: block sharding and the gather of match offsets to rank 0.

The reference has exactly one parallel strategy -- independent file blocks handed to a thread pool
(/root/reference/src/core/search_engine.cpp:104-172).  Every (block, alignment) chain is self-contained
(the (L-1)*W overlap bytes are part of the block, :227-230), so ranks take contiguous ranges of whole
blocks and scan them with NO data-path collective; only the result lists travel: an all-gather of the
counts followed by point-to-point sends to rank 0 (gather-v).  Rank-order concatenation is already
sorted by file offset because every match belongs to the block that contains its first byte.
"""


def shard_blocks(nblocks, rank, world):
    """-> (first_block, num_blocks) of this rank: contiguous, whole blocks, sizes differ by at most one."""
    b0 = rank * nblocks // world
    b1 = (rank + 1) * nblocks // world
    return b0, b1 - b0


def shard_bytes(file_size, block_size, overlap, rank, world):
    """-> (first_block, num_blocks, lo, hi): the byte range [lo, hi) of the file this rank must hold."""
    nblocks = (file_size + block_size - 1) // block_size
    b0, n = shard_blocks(nblocks, rank, world)
    lo = b0 * block_size
    hi = min(file_size, (b0 + n) * block_size + overlap) if n else lo
    return b0, n, lo, hi


def gather_offsets(dist, torch, offsets, rank, world, group=None):
    """Gather-v of per-rank int64 offset tensors to rank 0 (NCCL on GPUs, gloo in the CPU tests).
    Returns the concatenated tensor on rank 0 and None elsewhere."""
    n = torch.tensor([offsets.numel()], dtype=torch.int64, device=offsets.device)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    counts = [int(c.item()) for c in counts]
    if rank == 0:
        parts, reqs = [offsets], []
        for r in range(1, world):
            buf = torch.empty(counts[r], dtype=torch.int64, device=offsets.device)
            parts.append(buf)
            if counts[r]:
                reqs.append(dist.irecv(buf, src=r, group=group))
        for q in reqs:
            q.wait()
        return torch.cat(parts) if world > 1 else offsets
    if offsets.numel():
        dist.send(offsets, dst=0, group=group)
    return None


class PackedGather:
    """One NCCL collective per step instead of (all-gather of counts + send/recv) per search.

    Every rank packs the offset lists of all searches of a step into one fixed-size int64 buffer
    ``[n_0, ..., n_{k-1}, offsets_0 ..., offsets_{k-1} ...]`` and a single ``gather`` moves the buffers to
    rank 0.  A rank whose lists do not fit sends the remainder point-to-point; rank 0 learns that from
    the gathered header, so no extra collective is needed in the common (sparse) case.
    """

    def __init__(self, dist, torch, rank, world, nlists, capacity=8192, device="cuda", group=None):
        self.dist, self.torch, self.rank, self.world = dist, torch, rank, world
        self.nlists, self.cap, self.group = nlists, int(capacity), group
        self.buf = torch.zeros(self.cap, dtype=torch.int64, device=device)
        self.recv = [torch.zeros(self.cap, dtype=torch.int64, device=device) for _ in range(world)] if rank == 0 else None

    def __call__(self, lists):
        """lists: ``nlists`` int64 tensors on this rank.  Rank 0 gets ``nlists`` concatenated tensors (rank
        order == file order), other ranks get None."""
        torch, dist = self.torch, self.dist
        assert len(lists) == self.nlists
        room = self.cap - self.nlists
        at, spill = self.nlists, []
        header = []
        for t in lists:
            n = int(t.numel())
            header.append(n)
            fit = min(n, max(0, room - (at - self.nlists)))
            if fit:
                self.buf[at:at + fit] = t[:fit]
            at += fit
            if fit < n:
                spill.append(t[fit:])
        self.buf[: self.nlists] = torch.tensor(header, dtype=torch.int64).to(self.buf.device, non_blocking=True)
        dist.gather(self.buf, self.recv, dst=0, group=self.group)
        if self.rank != 0:
            for t in spill:
                dist.send(t.contiguous(), dst=0, group=self.group)
            return None
        heads = torch.stack([r[: self.nlists] for r in self.recv]).cpu().tolist()     # the only host read of a step
        out = [[] for _ in range(self.nlists)]
        for r in range(self.world):
            at, used = self.nlists, 0
            for k in range(self.nlists):
                n = heads[r][k]
                fit = min(n, max(0, room - used))
                piece = self.recv[r][at:at + fit].clone()
                at += fit
                used += fit
                if fit < n:
                    if r == 0:
                        rest = lists[k][fit:]
                    else:
                        rest = torch.empty(n - fit, dtype=torch.int64, device=self.buf.device)
                        dist.recv(rest, src=r, group=self.group)
                    piece = torch.cat([piece, rest])
                out[k].append(piece)
        return [torch.cat(p) for p in out]
