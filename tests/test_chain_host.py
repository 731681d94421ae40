"""ONE chain over slices (MonkeyMoore<Ty>::search on a buffer that is spread over GPUs) -- the host
side of the protocol, on CPU.  A slice hands on its MAP (entry phase -> exit phase); the entry phase of slice k is
the composition of the maps before it (mmg_chain_entry of the C-ABI, pure host arithmetic).  Here the oracle plays
the per-slice scan: its maps are found by brute force (one chain per entry phase), the composed entry phases must
reproduce the whole-buffer search exactly.  The GPU version of the same property is
tests/test_gpu_parity.py::test_chain_slices_equal_whole_search; the NCCL one is in tests/test_multigpu.py."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

KANA = "あいうえおかきくけこさしすせそたちつてとなにぬねのはひふへほまみむめもやゆよらりるれろわをゃっゅょ"
CASES = [
    ("8 monkey", 8, dict(keyword="monkey"), 256),
    ("8 abc low", 8, dict(keyword="abc"), 4),
    ("8 ab*de", 8, dict(keyword="ab*de", wildcard=ord("*")), 8),
    ("8 values", 8, dict(values=[10, 12, 15, 11, 30]), 64),
    ("16 mo*key*s", 16, dict(keyword="mo*key*s", wildcard=ord("*")), 16),
    ("16 kana", 16, dict(keyword="わたしたちは", char_seq=KANA), 64),
    ("16 MixedCase", 16, dict(keyword="MonKey"), 8),
]


def _data(bits, alphabet, n, seed):
    rng = np.random.default_rng(seed)
    return rng.integers(0, alphabet, n).astype(np.uint8 if bits == 8 else np.uint16)


def _slice_view(data, first, owned, tail):
    """elements a slice is given: what it owns plus keyword_len - 1 elements of its successor"""
    n = len(data)
    return data[first:min(n, first + owned + tail)] if first + owned < n else data[first:]


def _oracle_map(o, view, owned, jumps):
    return bytes(o.search_slice(view, e, owned)[2] - owned for e in range(jumps))


@pytest.mark.parametrize("name,bits,pat,alphabet", CASES, ids=[c[0] for c in CASES])
def test_composed_entry_phases_reproduce_the_whole_search(name, bits, pat, alphabet):
    import monkey_moore_b200 as mm
    from _oracle import Oracle
    prog = mm.Program(bits, **pat)
    o = Oracle(bits, **pat)
    jumps, tail = prog.max_jump, prog.keyword_len - 1
    assert 1 <= jumps <= 128
    n = 150_000 + 37
    data = _data(bits, alphabet, n, 7)
    want_pos, want_val = o.search(data)
    for slice_len in (4096 // (bits // 8), 40960, 65536 + 4096):
        firsts = list(range(0, n, slice_len))
        views = [_slice_view(data, f, min(slice_len, n - f), tail) for f in firsts]
        maps = [_oracle_map(o, v, min(slice_len, n - f), jumps) for f, v in zip(firsts[:-1], views[:-1])]
        assert all(max(m) < jumps for m in maps)
        pos, val = [], []
        for k, (f, v) in enumerate(zip(firsts, views)):
            entry = mm.chain_entry(maps[:k])
            p, x, _ = o.search_slice(v, entry, min(slice_len, n - f))
            pos.append(p + np.uint64(f))
            val.append(x)
        assert np.concatenate(pos).tolist() == want_pos.tolist(), (name, slice_len)
        assert np.concatenate(val).tolist() == want_val.tolist()


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import monkey_moore_b200 as mm
    from _oracle import Oracle
    results = {}
    for name, bits, pat, alphabet in CASES:
        prog, o = mm.Program(bits, **pat), Oracle(bits, **pat)
        jumps, tail = prog.max_jump, prog.keyword_len - 1
        n = 262144 + 4096 + 11                            # rank 1 owns a ragged tail
        data = _data(bits, alphabet, n, 11)               # (every rank could generate only its slice; the seed is shared)
        per = (n // world) // 4096 * 4096
        first = rank * per
        owned = per if rank < world - 1 else n - first
        view = _slice_view(data, first, owned, tail)
        last = rank == world - 1                          # nobody follows the last slice: its map is not needed
        mine = np.frombuffer((b"" if last else _oracle_map(o, view, owned, jumps)).ljust(128, b"\0"), np.uint8).copy()
        maps = [torch.zeros(128, dtype=torch.uint8) for _ in range(world)]
        dist.all_gather(maps, torch.from_numpy(mine))     # the one exchange of the path: 128 bytes per rank
        entry = mm.chain_entry([bytes(m.numpy().tolist()) for m in maps[:rank]])
        pos, val, _ = o.search_slice(view, entry, owned)
        pos = torch.from_numpy((pos + np.uint64(first)).astype(np.int64))
        counts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(counts, torch.tensor([len(pos)]))
        if rank == 0:
            parts = [pos]
            for r in range(1, world):
                buf = torch.zeros(int(counts[r]), dtype=torch.int64)
                if len(buf):
                    dist.recv(buf, src=r)
                parts.append(buf)
            whole, _ = o.search(data)
            results[name] = (torch.cat(parts).tolist() == whole.astype(np.int64).tolist(), len(whole), entry)
        elif len(pos):
            dist.send(pos, dst=0)
    if rank == 0:
        out.put(results)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_chain_search_over_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res and all(ok for ok, _, _ in res.values()), res
    assert sum(n for _, n, _ in res.values()) > 0
