"""Opt-in COMPLETE-MATCH mode (mmg_set_complete_matches): every matching window is reported, a superset of
what the reference's lossy skip chain visits.  Not reference behaviour, so its oracle -- the restatement's search loop
advancing by 1 -- is pinned differently: every reported window, searched ON ITS OWN by the unmodified reference, is a
match at position 0; every other window is not; and the chain's result is a subset.  GPU: the CUDA path in that mode
equals this oracle on all code paths (tiled, generic, evaluate-everything), engine blocks and chain slices included."""
import numpy as np
import pytest

from _cases import random_data, random_pattern
from _oracle import MMError as OracleError
from _oracle import Oracle, Ref


def pat_kwargs(p):
    return dict(keyword=p.get("keyword"), wildcard=p.get("wildcard", 0), char_seq=p.get("char_seq", ()),
                values=p.get("values"))


@pytest.fixture
def complete_oracle():
    old = Oracle.set_complete(True)
    yield
    Oracle.set_complete(old)


@pytest.mark.skipif(not Ref.available(), reason="needs the compiled reference (oracle/_ref/libmmref.so)")
@pytest.mark.parametrize("seed", range(3))
def test_complete_oracle_is_exactly_the_reference_match_predicate(seed):
    rng = np.random.default_rng(9000 + seed)
    checked = hits = 0
    for _ in range(25):
        bits = int(rng.choice([8, 16]))
        pat = random_pattern(rng, bits)
        try:
            o = Oracle(bits, **pat_kwargs(pat))
        except OracleError:
            continue
        data = random_data(rng, bits, int(rng.choice([40, 300, 1500])), pat)
        L = len(pat["values"]) if pat.get("values") is not None else len(pat["keyword"])
        chain_pos, _ = o.search(data)
        old = Oracle.set_complete(True)
        try:
            all_pos, _ = o.search(data)
        finally:
            Oracle.set_complete(old)
        assert set(chain_pos.tolist()) <= set(all_pos.tolist())
        assert np.all(all_pos[1:] > all_pos[:-1])
        want = set(all_pos.tolist())
        for s in range(0, max(len(data) - L + 1, 0)):
            pos, _ = Ref.search(bits, data[s:s + L], **pat_kwargs(pat))      # the window alone: a match iff position 0 comes back
            assert (len(pos) == 1 and int(pos[0]) == 0) == (s in want), (pat, bits, s)
            checked += 1
        hits += len(all_pos)
    assert checked > 1000 and hits > 0


def test_library_exports_the_switch():
    import monkey_moore_b200 as mm
    assert hasattr(mm.lib(), "mmg_set_complete_matches")


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(3))
def test_gpu_complete_mode_equals_the_oracle(gpu, complete_oracle, seed):
    rng = np.random.default_rng(9100 + seed)
    hits = extra = 0
    old = gpu.set_complete_matches(True)
    try:
        for it in range(40):
            bits = int(rng.choice([8, 16]))
            pat = random_pattern(rng, bits)
            try:
                o = Oracle(bits, **pat_kwargs(pat))
            except OracleError:
                continue
            W = bits // 8
            n = int(rng.choice([0, 5, 100, 4096, 20000, 300007]))
            data = random_data(rng, bits, n, pat)
            prog = gpu.Program(bits, **pat_kwargs(pat))
            want_pos, want_val = o.search(data)
            for override in (0, 1, 2):
                prev = gpu.set_path_override(override)
                try:
                    res = prog.search(data)
                finally:
                    gpu.set_path_override(prev)
                off, val = res.arrays()
                res.close()
                assert off.tolist() == want_pos.tolist(), (pat, bits, n, override)
                assert val.tolist() == want_val.tolist()
            # engine blocks (regular and irregular sizes, both endiannesses for 16 bit) and chain slices
            raw = data.view(np.uint8)
            for block in (4096, 65536, 1000, 47):
                if block < 64 and n > 5000:
                    continue
                be = bool(it & 1) and bits == 16
                eo, ev = o.engine(raw, block, big_endian=be, wrap32=False)
                res = prog.engine_scan(raw, block, big_endian=be)
                off, val = res.arrays()
                res.close()
                assert off.tolist() == eo.tolist() and val.tolist() == ev.tolist(), (pat, bits, n, block, be)
            if n >= 20000:
                parts = prog.search_sliced(data, 4096 // W * 3)
                got = np.concatenate([p.arrays()[0] for p in parts])
                for p in parts:
                    p.close()
                assert got.tolist() == want_pos.tolist()
            hits += len(want_pos)
            Oracle.set_complete(False)
            extra += len(want_pos) - len(o.search(data)[0])
            Oracle.set_complete(True)
    finally:
        gpu.set_complete_matches(old)
    assert hits > 0 and extra > 0          # the mode really reports matches the chain skips
