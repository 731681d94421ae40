"""Keywords longer than 128 elements.  The reference accepts any length (the GUI imposes no limit) and, in wildcard
mode, writes its skips through plain `char` (src/core/monkey_moore.cpp:253,269) -- values past 127 wrap on x86-64, the
ABI this parity targets.  CPU: the C restatement against the compiled reference for such keywords (this pins the
wrap); GPU: the product (per-chain kernels with the program's arrays in device memory) against the restatement."""
import numpy as np
import pytest

from _cases import HIRAGANA, random_data
from _oracle import Oracle, Ref


def long_patterns(rng):
    low = "abcdefghijklmnopqrstuvwxyz"
    out = []
    for L in (129, 130, 200, 300):
        out.append((8, dict(keyword="".join(rng.choice(list(low[:4])) for _ in range(L)))))
        kw = [rng.choice(list(low[:3])) for _ in range(L)]
        for _ in range(L // 8):
            kw[int(rng.integers(0, L))] = "*"
        out.append((8, dict(keyword="".join(kw), wildcard=ord("*"))))
        kw2 = list(kw)
        kw2[0] = "*"
        kw2[1] = "*"
        out.append((16, dict(keyword="".join(kw2), wildcard=ord("*"))))
        out.append((16, dict(keyword="".join(rng.choice(list(HIRAGANA[:6])) for _ in range(L)), char_seq=HIRAGANA)))
        out.append((8, dict(keyword="".join(rng.choice(list("aAbB")) for _ in range(L)))))          # mixed case -> wildcard mode
        out.append((8, dict(values=[int(rng.integers(0, 40)) for _ in range(L)])))
    return out


def kwargs(p):
    return dict(keyword=p.get("keyword"), wildcard=p.get("wildcard", 0), char_seq=p.get("char_seq", ()), values=p.get("values"))


@pytest.mark.skipif(not Ref.available(), reason="needs the compiled reference (oracle/_ref/libmmref.so)")
def test_restatement_equals_reference_for_long_keywords():
    rng = np.random.default_rng(31)
    hits = 0
    for bits, pat in long_patterns(rng):
        o = Oracle(bits, **kwargs(pat))
        for n in (100, 1000, 6000):
            data = random_data(rng, bits, n, pat)
            pos, vals = o.search(data)
            rpos, rmaps = Ref.search(bits, data, **kwargs(pat))
            assert pos.tolist() == rpos.tolist(), (bits, pat.get("keyword", "values")[:20], n)
            assert [o.table(int(v[0]), int(v[1])) for v in vals] == rmaps
            hits += len(pos)
    assert hits > 0


@pytest.mark.gpu
def test_gpu_long_keywords(gpu):
    rng = np.random.default_rng(32)
    hits = 0
    for bits, pat in long_patterns(rng):
        o = Oracle(bits, **kwargs(pat))
        prog = gpu.Program(bits, **kwargs(pat))
        assert prog.keyword_len > 128
        for n in (50, 1000, 40000):
            data = random_data(rng, bits, n, pat)
            want_pos, want_val = o.search(data)
            res = prog.search(data)
            off, val = res.arrays()
            assert res.stats()["fast_path"] == 0
            res.close()
            assert off.tolist() == want_pos.tolist(), (bits, n)
            assert [prog.table(int(v[0]), int(v[1])) for v in val] == [o.table(int(v[0]), int(v[1])) for v in want_val]
            raw = data.view(np.uint8)
            for block in (4096, 1000):
                be = bits == 16 and block == 4096
                eo, ev = o.engine(raw, block, big_endian=be, wrap32=False)
                r = prog.engine_scan(raw, block, big_endian=be)
                off, val = r.arrays()
                r.close()
                assert off.tolist() == eo.tolist() and val.tolist() == ev.tolist(), (bits, n, block)
            hits += len(want_pos)
        with pytest.raises(gpu.MMError):
            prog.chain_begin(np.zeros(8192, np.uint8 if bits == 8 else np.uint16), 8192, 0)
    assert hits > 0
