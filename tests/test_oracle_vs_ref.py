"""Differential pinning of the oracle against the UNMODIFIED reference compiled in place (oracle/_ref).
Skipped only when libmmref.so is absent (it cannot be rebuilt without /root/reference)."""
import os
import tempfile

import numpy as np
import pytest

from _cases import random_data, random_pattern
from _oracle import MMError, Oracle, Ref

pytestmark = pytest.mark.skipif(not Ref.available(), reason="oracle/_ref/libmmref.so not built")


def kw(p):
    return dict(keyword=p.get("keyword"), wildcard=p.get("wildcard", 0), char_seq=p.get("char_seq", ()),
                values=p.get("values"))


@pytest.mark.parametrize("seed", range(4))
def test_search_matches_reference(seed):
    rng = np.random.default_rng(seed)
    nonempty = 0
    for _ in range(400):
        bits = int(rng.choice([8, 16]))
        pat = random_pattern(rng, bits)
        try:
            o = Oracle(bits, **kw(pat))
        except MMError as e:
            if "non-terminating" not in str(e):
                assert not Ref.compile_ok(bits, **kw(pat))   # the reference throws for the same input
            continue
        assert Ref.compile_ok(bits, **kw(pat))
        data = random_data(rng, bits, int(rng.choice([0, 1, 5, 30, 200, 3000])), pat)
        pos, vals = o.search(data)
        rpos, rmaps = Ref.search(bits, data, **kw(pat))
        assert pos.tolist() == rpos.tolist(), pat
        assert [o.table(int(v[0]), int(v[1])) for v in vals] == rmaps, pat
        nonempty += len(pos) > 0
    assert nonempty > 50


@pytest.mark.parametrize("seed", range(2))
def test_engine_matches_reference(seed):
    rng = np.random.default_rng(100 + seed)
    path = os.path.join(tempfile.mkdtemp(), "blob.bin")
    nonempty = 0
    for _ in range(250):
        bits = int(rng.choice([8, 16]))
        pat = random_pattern(rng, bits)
        try:
            o = Oracle(bits, **kw(pat))
        except MMError:
            continue
        data = random_data(rng, bits, int(rng.choice([0, 1, 7, 64, 301, 2000])), pat)
        fb = np.ascontiguousarray(data).view(np.uint8)
        if rng.random() < 0.3 and len(fb):
            fb = np.ascontiguousarray(fb[:-1])
        fb.tofile(path)
        be = bool(rng.random() < 0.4)
        block = int(rng.choice([1, 2, 3, 5, 8, 16, 23, 29, 47, 58, 64, 100, 128, 512, 4096]))
        off, vals = o.engine(fb, block, big_endian=be)
        r = Ref.engine(bits, path, big_endian=be, threads=int(rng.choice([1, 3])), block=block, **kw(pat))
        assert off.tolist() == r["offsets"].tolist(), (pat, bits, block, be)
        assert [o.table(int(v[0]), int(v[1])) for v in vals] == r["maps"]
        nonempty += len(off) > 0
    assert nonempty > 30
