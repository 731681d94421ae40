"""Multi-GPU (needs >= 2 devices; skipped otherwise): every rank scans its block range on its own GPU,
the library gathers the match lists to rank 0 over NCCL, and the result equals the one-GPU scan and the
oracle -- for a sparse workload, for a dense one that overflows the packed buffer (spill path), and for a
dense one with more than 10^7 matches whose lists are released right after the gather call (the library
must keep them alive behind its own sends) while a collective of ANOTHER communicator follows at once.
Also: MonkeyMoore<Ty>::search as ONE chain over a buffer spread over the ranks (mmg_comm_search)."""
import dataclasses
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
    import faulthandler
    faulthandler.dump_traceback_later(240, exit=True)      # a hung collective must not hang the suite: show where, then die
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import monkey_moore_b200 as mm
    import monkey_moore_b200.workloads as wl
    from _oracle import Oracle, compose, digest
    from monkey_moore_b200.distributed import shard_bytes

    def bcast(raw):
        t = torch.tensor(list(raw), dtype=torch.uint8, device="cuda")
        dist.broadcast(t, src=0)
        return bytes(t.cpu().tolist())

    def oracle_of(w, s):
        pat = s.pattern
        return Oracle(w.bits, keyword=pat.get("keyword"), wildcard=pat.get("wildcard", 0),
                      char_seq=pat.get("char_seq", ()), values=pat.get("values"))

    comm = mm.Comm(rank, world, bcast, capacity=256)
    ok = {}
    # ---- small cases: full lists against the in-memory oracle and the one-GPU scan
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
            whole = wl.host_blob(w)
            for p, s, (off, val) in zip(progs, w.searches, got):
                exp, expv = oracle_of(w, s).engine(whole, w.block_size, big_endian=s.big_endian, wrap32=False)
                one = p.engine_scan(whole, w.block_size, big_endian=s.big_endian)
                o1, v1 = one.arrays()
                ok[key + ":" + s.name] = (off.tolist() == exp.tolist() and val.tolist() == expv.tolist()
                                          and o1.tolist() == exp.tolist(), len(exp))

    # ---- ONE chain over the whole buffer (MonkeyMoore<Ty>::search), slices spread over the ranks: the slice maps are
    # exchanged (all-gather of 128 bytes per rank), every rank finishes with its composed entry phase
    for key, size in (("cfg5", 8 << 20), ("cfg2", 32 << 20), ("cfg1", 16 << 20)):
        w = wl.WORKLOADS[key].scaled(size)
        s = w.searches[0]
        prog = mm.Program(w.bits, **s.pattern)
        W = w.bits // 8
        n = w.size // W
        per = (n // world) * W // 4096 * 4096 // W
        first = rank * per
        owned = per if rank < world - 1 else n - first
        avail = owned if rank == world - 1 else owned + prog.keyword_len - 1
        blob = wl.device_blob(w, first_byte=first * W, nbytes=avail * W, total_size=w.size)
        res = comm.search(prog, blob, owned, first)
        got = comm.gather([res], fetch=True)
        res.close()
        if rank == 0:
            whole = wl.host_blob(w)
            exp, expv = oracle_of(w, s).search(whole.view(np.uint16) if W == 2 else whole)
            off, val = got[0]
            ok["chain:" + key] = (off.tolist() == exp.tolist() and val.tolist() == expv.tolist(), len(exp))

    # ---- dense case, > 10^7 matches: 4-symbol alphabet, lists freed immediately, two gathers back to back, then a
    # collective of torch's communicator while this rank's spill sends may still be in flight
    w = dataclasses.replace(wl.WORKLOADS["cfg5"], size=768 << 20, byte_mask=0x03)
    s = w.searches[0]
    prog = mm.Program(w.bits, **s.pattern)
    overlap = (prog.keyword_len - 1) * (w.bits // 8)
    b0, nb, lo, hi = shard_bytes(w.size, w.block_size, overlap, rank, world)
    blob = wl.device_blob(w, first_byte=lo, nbytes=hi - lo, total_size=w.size)
    gathered = []
    for _ in range(2):
        res = prog.engine_scan(blob, w.block_size, file_size=w.size, first_block=b0, num_blocks=nb)
        g = comm.gather([res], lazy=True)
        res.close()                                      # released while the gather may still read it
        scratch = torch.full((64 << 20,), 0xFF, dtype=torch.uint8, device="cuda")     # recycle freed memory eagerly
        del scratch
        gathered.append(g)
    flag = torch.ones(1, device="cuda")
    dist.all_reduce(flag)                                # another communicator's collective right behind the gathers
    o = oracle_of(w, s)
    r = o.engine_synth(w.seed, w.byte_mask, w.size, w.block_size, b0, nb, False, [], threads=max(1, (os.cpu_count() or 2) // world))
    t = torch.tensor([x - (1 << 64) if x >= (1 << 63) else x for x in (r["count"], r["s0"], r["s1"])], dtype=torch.int64, device="cuda")
    parts = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(parts, t)
    if rank == 0:
        want = compose([tuple(int(x) & ((1 << 64) - 1) for x in p.tolist()) for p in parts])
        off, val = gathered[1].fetch()[0]
        good = digest(off, val) == want and bool(np.all(off[1:] > off[:-1])) and gathered[0].counts()[0] == want[0]
        ok["dense4:" + s.name] = (good, int(want[0]))
        for g in gathered:
            g.close()
    comm.wait()
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
    res = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert res and all(good for good, _ in res.values()), res
    assert res["cfg5:8bit-abc-dense"][1] > 1000      # dense enough to exercise the spill path
    assert res["dense4:8bit-abc-dense"][1] > 10_000_000
    assert res["chain:cfg5"][1] > 1000 and res["chain:cfg2"][1] > 0
