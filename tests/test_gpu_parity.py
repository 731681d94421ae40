"""GPU parity: the CUDA path (through the C-ABI) against the oracle on the same seeded inputs.

Bar: bit-exact -- same match offsets, same order, same inferred tables."""
import numpy as np
import pytest

from _cases import (ENGINE8_BLOCKS, ENGINE8_FILE, ENGINE8_OFFSETS, ENGINE16_BLOCKS_BE, ENGINE16_BLOCKS_LE,
                    ENGINE16_FILE, ENGINE16_OFFSETS, random_data, random_pattern, ref_kats)
from _oracle import MMError as OracleError
from _oracle import Oracle

pytestmark = pytest.mark.gpu


def pat_kwargs(p):
    return dict(keyword=p.get("keyword"), wildcard=p.get("wildcard", 0), char_seq=p.get("char_seq", ()),
                values=p.get("values"))


def check_search(mm, bits, pat, data, override=0):
    o = Oracle(bits, **pat_kwargs(pat))
    prog = mm.Program(bits, **pat_kwargs(pat))
    old = mm.set_path_override(override)
    try:
        res = prog.search(data)
    finally:
        mm.set_path_override(old)
    off, val = res.arrays()
    opos, ovals = o.search(data)
    assert off.tolist() == opos.tolist(), (pat, bits, len(data), off[:8], opos[:8])
    assert [prog.table(int(v[0]), int(v[1])) for v in val] == [o.table(int(v[0]), int(v[1])) for v in ovals]
    return len(off)


def check_engine(mm, bits, pat, file_bytes, block, big_endian, override=0):
    o = Oracle(bits, **pat_kwargs(pat))
    prog = mm.Program(bits, **pat_kwargs(pat))
    old = mm.set_path_override(override)
    try:
        res = prog.engine_scan(file_bytes, block, big_endian=big_endian)
    finally:
        mm.set_path_override(old)
    off, val = res.arrays()
    ooff, ovals = o.engine(file_bytes, block, big_endian=big_endian, wrap32=False)
    assert off.tolist() == ooff.tolist(), (pat, bits, len(file_bytes), block, big_endian, off[:8], ooff[:8])
    assert [prog.table(int(v[0]), int(v[1])) for v in val] == [o.table(int(v[0]), int(v[1])) for v in ovals]
    return len(off)


@pytest.mark.parametrize("kat", ref_kats(), ids=lambda k: k["name"])
@pytest.mark.parametrize("override", [0, 1, 2])
def test_reference_known_answers(gpu, kat, override):
    """/root/reference/tests/test_monkey_moore.cpp -- offsets and full value tables."""
    prog = gpu.Program(kat["bits"], **pat_kwargs(kat))
    old = gpu.set_path_override(override)
    try:
        res = prog.search(kat["data"])
    finally:
        gpu.set_path_override(old)
    assert res.offsets.tolist() == kat["pos"]
    if kat["maps"] is not None:
        assert res.tables() == kat["maps"]


@pytest.mark.parametrize("override", [0, 1, 2])
def test_reference_engine_known_answers(gpu, override):
    """/root/reference/tests/test_search_engine.cpp:26-158 -- tiny files, pathological block sizes, BE."""
    old = gpu.set_path_override(override)
    try:
        cases = [(8, ENGINE8_FILE, ENGINE8_OFFSETS, ENGINE8_BLOCKS, False),
                 (16, ENGINE16_FILE, ENGINE16_OFFSETS, ENGINE16_BLOCKS_LE, False),
                 (16, ENGINE16_FILE.byteswap(), ENGINE16_OFFSETS, ENGINE16_BLOCKS_BE, True)]
        for bits, file, offs, blocks, be in cases:
            fb = np.ascontiguousarray(file).view(np.uint8)
            prog = gpu.Program(bits, keyword="text", wildcard=ord("*"))
            for b in blocks:
                assert prog.engine_scan(fb, b, big_endian=be).offsets.tolist() == offs, (bits, b, be)
    finally:
        gpu.set_path_override(old)


@pytest.mark.parametrize("seed", range(6))
def test_search_fuzz_small(gpu, seed):
    """Randomised patterns x data styles, every path (tiled, generic, evaluate-everything)."""
    rng = np.random.default_rng(1000 + seed)
    hits = 0
    for _ in range(60):
        bits = int(rng.choice([8, 16]))
        pat = random_pattern(rng, bits)
        try:
            Oracle(bits, **pat_kwargs(pat))
        except OracleError:
            continue
        n = int(rng.choice([0, 1, 3, 17, 100, 511, 4096, 5000, 20000]))
        data = random_data(rng, bits, n, pat)
        for override in (0, 1, 2):
            hits += check_search(gpu, bits, pat, data, override)
    assert hits > 0


@pytest.mark.parametrize("seed", range(4))
def test_search_fuzz_large(gpu, seed):
    """Longer single chains: many sub-tiles, phase maps composed across tiles."""
    rng = np.random.default_rng(2000 + seed)
    hits = 0
    for _ in range(12):
        bits = int(rng.choice([8, 16]))
        pat = random_pattern(rng, bits)
        try:
            Oracle(bits, **pat_kwargs(pat))
        except OracleError:
            continue
        n = int(rng.choice([70000, 262144, 1000003]))
        data = random_data(rng, bits, n, pat)
        hits += check_search(gpu, bits, pat, data)
    assert hits > 0


@pytest.mark.parametrize("seed", range(4))
def test_engine_fuzz(gpu, seed):
    """Block decomposition: regular blocks (tiled path), irregular ones (generic path), LE/BE, odd sizes."""
    rng = np.random.default_rng(3000 + seed)
    hits = 0
    for _ in range(40):
        bits = int(rng.choice([8, 16]))
        pat = random_pattern(rng, bits)
        try:
            Oracle(bits, **pat_kwargs(pat))
        except OracleError:
            continue
        n = int(rng.choice([0, 5, 64, 301, 5000, 40000, 150001]))
        data = random_data(rng, bits, n, pat)
        fb = np.ascontiguousarray(data).view(np.uint8)
        if rng.random() < 0.3 and len(fb) > 0:
            fb = np.ascontiguousarray(fb[:-1])
        block = int(rng.choice([1, 3, 8, 23, 47, 100, 512, 4096, 4096, 8192, 16384, 65536, 524288]))
        be = bool(rng.random() < 0.4)
        hits += check_engine(gpu, bits, pat, fb, block, be)
    assert hits > 0


def test_sharded_scan_equals_whole(gpu):
    """Scanning block ranges separately (one per rank) and concatenating == scanning the file once."""
    rng = np.random.default_rng(7)
    for bits, pat in [(8, dict(keyword="abc")), (16, dict(keyword="mo*key*s", wildcard=ord("*")))]:
        data = random_data(rng, bits, 300000, pat)
        fb = np.ascontiguousarray(data).view(np.uint8)
        block = 16384
        prog = gpu.Program(bits, **pat_kwargs(pat))
        whole = prog.engine_scan(fb, block).offsets
        nblocks = (len(fb) + block - 1) // block
        overlap = (prog.keyword_len - 1) * (bits // 8)
        parts = []
        for r in range(3):
            b0, b1 = r * nblocks // 3, (r + 1) * nblocks // 3
            lo, hi = b0 * block, min(len(fb), b1 * block + overlap)
            parts.append(prog.engine_scan(np.ascontiguousarray(fb[lo:hi]), block, file_size=len(fb),
                                          first_block=b0, num_blocks=b1 - b0).offsets)
        assert np.concatenate(parts).tolist() == whole.tolist()


@pytest.mark.parametrize("seed", range(3))
def test_engine_big_blocks(gpu, seed):
    """Engine blocks of more than 128 sub-tiles (the GUI's 8 MiB default, src/gui/monkey_prefs.cpp:26) are resolved in
    segments: segment maps, phase prefix along the block, then the usual replay -- same results as one chain per block."""
    rng = np.random.default_rng(4000 + seed)
    hits = 0
    for _ in range(6):
        bits = int(rng.choice([8, 16]))
        pat = random_pattern(rng, bits)
        try:
            Oracle(bits, **pat_kwargs(pat))
        except OracleError:
            continue
        n = int(rng.choice([1500000, 2500001, 5000000]))
        data = random_data(rng, bits, n, pat)
        fb = np.ascontiguousarray(data).view(np.uint8)
        block = int(rng.choice([1 << 20, 2 << 20, 8 << 20, 528384]))
        be = bool(rng.random() < 0.4)
        hits += check_engine(gpu, bits, pat, fb, block, be)
    assert hits > 0


def test_sparse_resolve_kernel_and_its_fallback(gpu):
    """The second scan of a pattern whose first scan left few events per block is resolved by the warp-per-block kernel;
    a later scan over dense data makes that kernel hand over to the general one.  Results stay bit-exact throughout."""
    rng = np.random.default_rng(77)
    sparse = rng.integers(0, 256, size=3 << 20, dtype=np.uint8)
    for v in (10, 100000, 2000000, (3 << 20) - 9):          # planted 8-bit matches, one right before the end
        sparse[v:v + 6] = np.frombuffer(b"monkey", dtype=np.uint8) + 3
    sparse[524288 - 3:524288 + 3] = np.frombuffer(b"monkey", dtype=np.uint8)

    def dense16(be):        # 16-bit elements from an 8-symbol alphabet: every ninth difference is +1
        e = rng.integers(100, 108, size=3 << 19, dtype=np.uint16)
        return np.ascontiguousarray(e.byteswap() if be else e).view(np.uint8)

    def planted16(be):      # sparse 16-bit data with shifted copies of "acegh", aligned and odd, one across a block edge
        b = sparse.copy()
        for v in (64, 70001, 524288 - 4, 1500000, 2999990):
            e = (np.array([0x4100, 0x4102, 0x4104, 0x4106, 0x4107], dtype=np.uint16) + np.uint16(v % 97))
            b[v:v + 10] = np.ascontiguousarray(e.byteswap() if be else e).view(np.uint8)
        return b

    cases = [(8, dict(keyword="monkey"), False, sparse, rng.integers(0, 4, size=3 << 20, dtype=np.uint8), False),
             (16, dict(keyword="mo*key*s", wildcard=ord("*")), False, sparse, dense16(False), False),
             (16, dict(keyword="acegh"), False, planted16(False), dense16(False), True),
             (16, dict(keyword="acegh"), True, planted16(True), dense16(True), True)]
    for bits, pat, be, few, many, expect_handover in cases:
        prog = gpu.Program(bits, **pat)
        o = Oracle(bits, **pat_kwargs(pat))
        kinds, counts = [], []
        for blob in (few, few, many, few, few):
            res = prog.engine_scan(blob, 524288, big_endian=be)
            off, val = res.arrays()
            kinds.append(res.stats()["resolve_kind"])
            counts.append(len(off))
            exp, expv = o.engine(blob, 524288, big_endian=be, wrap32=False)
            assert off.tolist() == exp.tolist() and val.tolist() == expv.tolist(), (bits, pat, be, kinds)
        assert kinds[0] == 0, kinds                       # no hint yet: general kernel
        if expect_handover:                               # sparse, handed over on dense data, general, sparse again
            assert kinds == [0, 1, 2, 0, 1], (kinds, counts)
            assert counts[0] >= 2, counts


def chain_lists(parts):
    off = np.concatenate([p.arrays()[0] for p in parts]) if parts else np.zeros(0, np.uint64)
    val = np.concatenate([p.arrays()[1] for p in parts]) if parts else np.zeros((0, 2), np.uint32)
    for p in parts:
        p.close()
    return off, val


@pytest.mark.parametrize("seed", range(4))
def test_chain_slices_equal_whole_search(gpu, seed):
    """ONE chain over a buffer cut into slices (mmg_chain_begin / _map / _entry / _finish): the slices'
    lists, concatenated, are MonkeyMoore<Ty>::search over the whole buffer -- from host and from device memory, for
    slices of one sub-tile, of exactly one 128-sub-tile segment (a trailing overlap segment appears), of several
    segments, and with other scans using the workspace between the two halves (the slice then re-runs itself)."""
    import torch
    rng = np.random.default_rng(7000 + seed)
    hits = cases = 0
    for it in range(10):
        bits = int(rng.choice([8, 16]))
        pat = random_pattern(rng, bits)
        try:
            o = Oracle(bits, **pat_kwargs(pat))
        except OracleError:
            continue
        W = bits // 8
        n = int(rng.choice([4096 // W * 3 + 5, 300_007, 1_200_011, 3_000_000]))
        data = random_data(rng, bits, n, pat)
        want_pos, want_val = o.search(data)
        prog = gpu.Program(bits, **pat_kwargs(pat))
        dev = torch.from_numpy(data.view(np.uint8)).cuda()
        for sub_tiles in (1, 128, 130, 512):
            slice_len = sub_tiles * 4096 // W
            if slice_len * 400 < n:
                continue                      # (hundreds of one-sub-tile slices of a large buffer: nothing new, just slow)
            src = (dev.data_ptr(), dev.numel()) if (it + sub_tiles) % 2 else data
            off, val = chain_lists(prog.search_sliced(src, slice_len))
            assert off.tolist() == want_pos.tolist(), (pat, bits, n, sub_tiles)
            assert [prog.table(int(v[0]), int(v[1])) for v in val] == [o.table(int(v[0]), int(v[1])) for v in want_val]
            cases += 1
        # the workspace is used by other scans between the halves
        slice_len = 64 * 4096 // W
        tail = prog.keyword_len - 1
        slices = []
        first = 0
        while first < n:
            owned = min(slice_len, n - first)
            if n - (first + owned) <= tail:
                owned = n - first
            view = data[first:] if first + owned >= n else data[first:first + owned + tail]
            slices.append(prog.chain_begin(view, owned, first))
            prog.search(data[: 50_000 // W]).close()
            first += owned
        maps = [s.map() for s in slices]
        assert all(len(m) == prog.max_jump for m in maps)
        prog.engine_scan(data.view(np.uint8)[: 1 << 20], 65536).close()
        parts = [s.finish(gpu.chain_entry(maps[:k])) for k, s in enumerate(slices)]
        off, val = chain_lists(parts)
        assert off.tolist() == want_pos.tolist(), (pat, bits, n, "interleaved")
        hits += len(want_pos)
    assert hits > 0 and cases > 0


def test_chain_slice_argument_errors(gpu):
    prog = gpu.Program(8, keyword="monkey")
    data = np.zeros(10000, np.uint8)
    with pytest.raises(gpu.MMError):
        prog.chain_begin(data, 5000, 0)                  # continued slice that does not own a multiple of 4096 bytes
    with pytest.raises(gpu.MMError):
        prog.chain_begin(data[:4098], 4096, 0)           # continued, but fewer than keyword_len - 1 elements behind it
    s = prog.chain_begin(data[:8192 + 5], 8192, 0)
    s.close()                                            # abandoned before its map was read
    s = prog.chain_begin(data, 10000, 0)
    with pytest.raises(gpu.MMError):
        s.finish(200)                                    # entry phase outside the jump range
