"""world_size-2 (gloo, CPU) test of the multi-GPU host logic: block sharding + gather-v to rank 0.
The per-rank scan is played by the oracle here (no GPU in this container); the GPU version of the
same property is tests/test_gpu_parity.py::test_sharded_scan_equals_whole."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import monkey_moore_b200.workloads as wl
    from _oracle import Oracle
    from monkey_moore_b200.distributed import PackedGather, gather_offsets, shard_bytes

    w = wl.WORKLOADS["cfg2"].scaled(3 << 20)
    w.block_size = 65536
    results = {}
    mine_lists, whole_lists = [], []
    for s in w.searches:
        o = Oracle(w.bits, keyword=s.pattern["keyword"], wildcard=s.pattern["wildcard"])
        overlap = (len(s.pattern["keyword"]) - 1) * 2
        b0, nb, lo, hi = shard_bytes(w.size, w.block_size, overlap, rank, world)
        mine = wl.host_blob(w, first_byte=lo, nbytes=hi - lo)          # each rank materialises only its slice
        off, _ = o.engine(mine, w.block_size, big_endian=s.big_endian, wrap32=False)
        # a match belongs to the block holding its first byte: drop what the next rank owns
        off = off[off < np.uint64(nb * w.block_size)] + np.uint64(lo)
        g = gather_offsets(dist, torch, torch.from_numpy(off.astype(np.int64)), rank, world)
        mine_lists.append(torch.from_numpy(off.astype(np.int64)))
        if rank == 0:
            whole, _ = o.engine(wl.host_blob(w), w.block_size, big_endian=s.big_endian, wrap32=False)
            whole_lists.append(whole.astype(np.int64).tolist())
            results[s.name] = (g.numpy().tolist() == whole.astype(np.int64).tolist(), len(whole))
    # the one-collective-per-step gather, once roomy and once so small that every list spills
    for cap in (8192, 12):
        pg = PackedGather(dist, torch, rank, world, len(mine_lists), capacity=cap, device="cpu")
        got = pg(mine_lists)
        if rank == 0:
            results["packed-%d" % cap] = ([g.tolist() for g in got] == whole_lists, sum(len(x) for x in whole_lists))
    if rank == 0:
        out.put(results)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_gather():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res and all(ok and n > 0 for ok, n in res.values()), res


def test_shard_arithmetic():
    sys.path.insert(0, ROOT)
    from monkey_moore_b200.distributed import shard_blocks, shard_bytes
    for nblocks in (0, 1, 7, 8, 1024, 1025):
        for world in (1, 2, 4, 8):
            parts = [shard_blocks(nblocks, r, world) for r in range(world)]
            assert sum(n for _, n in parts) == nblocks
            assert all(parts[i][0] + parts[i][1] == parts[i + 1][0] for i in range(world - 1))
    b0, n, lo, hi = shard_bytes(1000, 100, 14, 1, 2)
    assert (b0, n, lo, hi) == (5, 5, 500, 1000)
    b0, n, lo, hi = shard_bytes(1000, 100, 14, 0, 2)
    assert (b0, n, lo, hi) == (0, 5, 0, 514)
