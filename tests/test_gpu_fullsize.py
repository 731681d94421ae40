"""GPU parity at BASELINE.json's sizes (and beyond 4 GiB), through size-independent properties plus the
oracle where it finishes in seconds.

* cfg3 (8-bit value scan, 4 GiB): planted matches are found; scanning the blob as two block ranges equals
  scanning it whole; the first 1 GiB equals the oracle bit for bit.
* cfg5 (dense 3-char keyword over a 16-symbol blob): 1 GiB equals the oracle (millions of matches).
* > 4 GiB: 64-bit offsets -- matches planted beyond 4 GiB are reported at their true offsets (the reference
  wraps its block offsets at 32 bit, /root/reference/src/core/search_engine.cpp:241-242; documented deviation),
  checked against the oracle with 64-bit block arithmetic.
"""
import numpy as np
import pytest

from _oracle import Oracle

pytestmark = pytest.mark.gpu

GiB = 1 << 30


def _oracle_for(w, s):
    p = s.pattern
    return Oracle(w.bits, keyword=p.get("keyword"), wildcard=p.get("wildcard", 0), char_seq=p.get("char_seq", ()),
                  values=p.get("values"))


def test_cfg3_value_scan_4gib(gpu):
    import monkey_moore_b200.workloads as wl
    w = wl.WORKLOADS["cfg3"]
    s = w.searches[0]
    prog = gpu.Program(w.bits, **s.pattern)
    blob = wl.device_blob(w)
    whole = prog.engine_scan(blob, w.block_size)
    off = whole.offsets
    assert whole.stats()["fast_path"] == 1
    # every planted match is a true match; the chain may skip a few, never invent one
    planted = sorted(o for o, _ in wl.planted_patches(w, w.size))
    assert len(planted) >= 1000
    found = np.isin(np.array(planted, dtype=np.uint64), off)
    assert found.mean() > 0.7, found.mean()   # the lossy skip chain legitimately drops some
    assert np.all(np.diff(off.astype(np.int64)) > 0)          # strictly ascending
    # two block ranges == whole
    nblocks = w.size // w.block_size
    overlap = (prog.keyword_len - 1)
    half = nblocks // 2
    a = prog.engine_scan(blob[: half * w.block_size + overlap], w.block_size, file_size=w.size, first_block=0, num_blocks=half).offsets
    b = prog.engine_scan(blob[half * w.block_size:], w.block_size, file_size=w.size, first_block=half, num_blocks=nblocks - half).offsets
    assert np.array_equal(np.concatenate([a, b]), off)
    # first GiB against the oracle (a prefix of whole blocks is scanned identically)
    o = _oracle_for(w, s)
    host = wl.host_blob(w, first_byte=0, nbytes=GiB + overlap)
    exp, _ = o.engine(host, w.block_size, wrap32=False)
    exp = exp[exp < GiB]
    assert np.array_equal(off[off < GiB], exp)


def test_cfg5_dense_1gib(gpu):
    import monkey_moore_b200.workloads as wl
    w = wl.WORKLOADS["cfg5"].scaled(GiB)
    s = w.searches[0]
    prog = gpu.Program(w.bits, **s.pattern)
    host = wl.host_blob(w)
    res = prog.engine_scan(wl.device_blob(w), w.block_size)
    off, val = res.arrays()
    exp, expv = _oracle_for(w, s).engine(host, w.block_size, wrap32=False)
    assert len(exp) > 1_000_000
    assert np.array_equal(off, exp) and np.array_equal(val, expv)


def test_offsets_beyond_4gib(gpu):
    import monkey_moore_b200.workloads as wl
    w = wl.WORKLOADS["cfg1"].scaled(4 * GiB + (96 << 20))
    w.n_planted = 4000
    s = w.searches[0]
    prog = gpu.Program(w.bits, **s.pattern)
    res = prog.engine_scan(wl.device_blob(w), w.block_size)
    off = res.offsets
    assert (off >= np.uint64(4 * GiB)).sum() > 10              # matches past the 32-bit boundary exist ...
    host = wl.host_blob(w, first_byte=4 * GiB - w.block_size, nbytes=w.size - (4 * GiB - w.block_size))
    exp, _ = _oracle_for(w, s).engine(host, w.block_size, wrap32=False)
    exp = exp + np.uint64(4 * GiB - w.block_size)             # ... at exactly the oracle's 64-bit offsets
    assert np.array_equal(off[off >= np.uint64(4 * GiB - w.block_size)], exp)


@pytest.mark.parametrize("key,size", [("cfg5", 200 << 20), ("cfg2", 160 << 20)])
def test_one_chain_over_hundreds_of_segments(gpu, key, size):
    """MonkeyMoore<Ty>::search on ONE large buffer: a single block of several hundred 128-sub-tile segments, whose
    phase prefix runs in two levels (k_rangemap, k_chainphase) -- in one call, and as two slices of ~200 segments
    each through mmg_chain_* (ranges of two segments).  Equal to the oracle's chain over the whole buffer."""
    import monkey_moore_b200.workloads as wl
    w = wl.WORKLOADS[key].scaled(size)
    s = w.searches[0]
    prog = gpu.Program(w.bits, **s.pattern)
    W = w.bits // 8
    blob = wl.device_blob(w)
    host = wl.host_blob(w)
    elems = host.view(np.uint16) if W == 2 else host
    exp, expv = _oracle_for(w, s).search(elems)
    res = prog.search(blob)
    off, val = res.arrays()
    assert res.stats()["fast_path"] == 1 and res.stats()["launches"] >= 5
    res.close()
    assert off.tolist() == exp.tolist() and val.tolist() == expv.tolist()
    n = len(elems)
    parts = prog.search_sliced(blob, (n // 2) * W // 4096 * 4096 // W)
    assert len(parts) == 2
    soff = np.concatenate([p.arrays()[0] for p in parts])
    sval = np.concatenate([p.arrays()[1] for p in parts])
    for p in parts:
        p.close()
    assert soff.tolist() == exp.tolist() and sval.tolist() == expv.tolist()
    assert len(exp) > (1000 if key == "cfg5" else 0)
