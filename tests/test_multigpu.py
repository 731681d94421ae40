"""Multi-GPU (needs >= 2 devices; skipped otherwise): every rank scans its block range on its own GPU,
the library gathers the match lists to rank 0 over NCCL, and the result equals the one-GPU scan and the
oracle -- both for a sparse workload and for a dense one that overflows the packed buffer (spill path)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import monkey_moore_b200 as mm
    import monkey_moore_b200.workloads as wl
    from monkey_moore_b200.distributed import shard_bytes

    def bcast(raw):
        t = torch.tensor(list(raw), dtype=torch.uint8, device="cuda")
        dist.broadcast(t, src=0)
        return bytes(t.cpu().tolist())

    comm = mm.Comm(rank, world, bcast, capacity=256)
    ok = {}
    for key, size in (("cfg2", 24 << 20), ("cfg5", 6 << 20)):
        w = wl.WORKLOADS[key].scaled(size)
        progs = [mm.Program(w.bits, **s.pattern) for s in w.searches]
        overlap = (max(p.keyword_len for p in progs) - 1) * (w.bits // 8)
        b0, nb, lo, hi = shard_bytes(w.size, w.block_size, overlap, rank, world)
        blob = wl.device_blob(w, first_byte=lo, nbytes=hi - lo, total_size=w.size)
        held = [p.engine_scan(blob, w.block_size, big_endian=s.big_endian, file_size=w.size, first_block=b0, num_blocks=nb)
                for p, s in zip(progs, w.searches)]
        got = comm.gather(held, fetch=True)
        if rank == 0:
            from _oracle import Oracle
            whole = wl.host_blob(w)
            for p, s, (off, val) in zip(progs, w.searches, got):
                pat = s.pattern
                o = Oracle(w.bits, keyword=pat.get("keyword"), wildcard=pat.get("wildcard", 0),
                           char_seq=pat.get("char_seq", ()), values=pat.get("values"))
                exp, expv = o.engine(whole, w.block_size, big_endian=s.big_endian, wrap32=False)
                one = p.engine_scan(whole, w.block_size, big_endian=s.big_endian)
                o1, v1 = one.arrays()
                ok[key + ":" + s.name] = (off.tolist() == exp.tolist() and val.tolist() == expv.tolist()
                                          and o1.tolist() == exp.tolist(), len(exp))
    if rank == 0:
        q.put(ok)
    dist.barrier()
    comm.close()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_scan_and_nccl_gather():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert res and all(good for good, _ in res.values()), res
    assert res["cfg5:8bit-abc-dense"][1] > 1000      # dense enough to exercise the spill path
