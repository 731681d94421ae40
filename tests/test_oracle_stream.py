"""oracle/mm_oracle_stream.c (the engine over a synthetic file that is regenerated block by block, with its
order-sensitive digest) against oracle/mm_oracle.c's in-memory engine: lists, digests, and the composition of the
digests of adjacent block ranges -- on small instances of all five BASELINE configurations.  CPU only."""
import numpy as np
import pytest

import monkey_moore_b200.workloads as wl
from _oracle import Oracle, compose, digest

CASES = [("cfg1", 2 << 20), ("cfg2", 3 << 20), ("cfg3", 2 << 20), ("cfg4", 3 << 20), ("cfg5", 1 << 20)]


def oracle_of(w, s):
    p = s.pattern
    return Oracle(w.bits, keyword=p.get("keyword"), wildcard=p.get("wildcard", 0), char_seq=p.get("char_seq", ()),
                  values=p.get("values"))


@pytest.mark.parametrize("key,size", CASES)
def test_streamed_engine_equals_in_memory_engine(key, size):
    w = wl.WORKLOADS[key].scaled(size)
    if w.generator != "splitmix":
        pytest.skip("the streamed oracle regenerates the counter-based blobs only (cfg1 uses the reference's mt19937 data)")
    blob = wl.host_blob(w)
    patches = wl.planted_patches(w, w.size)
    nb = (w.size + w.block_size - 1) // w.block_size
    for s in w.searches:
        o = oracle_of(w, s)
        exp, expv = o.engine(blob, w.block_size, big_endian=s.big_endian, wrap32=False)
        r = o.engine_synth(w.seed, w.byte_mask, w.size, w.block_size, big_endian=s.big_endian, patches=patches, threads=3,
                           want_list=True)
        assert r["offsets"].tolist() == exp.tolist() and r["values"].tolist() == expv.tolist()
        whole = digest(exp, expv)
        assert (r["count"], r["s0"], r["s1"]) == whole
        assert len(exp) > 0
        # digests of adjacent block ranges compose to the digest of the whole
        cut = max(1, nb // 3)
        parts = [o.engine_synth(w.seed, w.byte_mask, w.size, w.block_size, a, n, s.big_endian, patches, threads=2)
                 for a, n in ((0, cut), (cut, nb - cut))]
        assert compose([(p["count"], p["s0"], p["s1"]) for p in parts]) == whole


def test_digest_is_order_sensitive():
    off = np.array([5, 9, 100], dtype=np.uint64)
    val = np.array([[1, 0], [2, 0], [3, 0]], dtype=np.uint32)
    a = digest(off, val)
    b = digest(off[[1, 0, 2]], val[[1, 0, 2]])
    assert a[0] == b[0] and a[1] == b[1] and a[2] != b[2]
    assert digest(off, val, first_index=7) != a
