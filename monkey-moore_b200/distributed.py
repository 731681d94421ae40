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
