"""Distinct inferred tables on the device against the reference GUI's rule.

/root/reference/src/gui/monkey_frame.cpp:1236-1245 lists a result iff no EARLIER result has an equal values map;
the expected index list below is that rule applied to the tables of the whole match list on the host."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HIRAGANA = "あいうえおかきくけこさしすせそたちつてとなにぬねのはひふへほまみむめもやゆよらりるれろわをゃっゅょ"


def gui_unique(prog, val):
    """first index of every distinct table, in list order (the GUI's `unique` vector)"""
    seen, first, cache = set(), [], {}
    for i, v in enumerate(val):
        key = (int(v[0]), int(v[1]))
        if key not in cache:
            cache[key] = tuple(sorted(prog.table(*key).items()))
        t = cache[key]
        if t not in seen:
            seen.add(t)
            first.append(i)
    return first


def low_entropy(n, symbols, seed, bits=8):
    rng = np.random.default_rng(seed)
    if bits == 8:
        return rng.integers(0, symbols, size=n, dtype=np.uint8)
    return rng.integers(0, symbols, size=n // 2, dtype=np.uint16).view(np.uint8)


CASES = [
    # name, bits, pattern, data, expected path
    ("8bit-simple-dense", 8, dict(keyword="abc"), low_entropy(1 << 20, 16, 1), "direct"),
    ("8bit-mixed-case", 8, dict(keyword="aBc"), low_entropy(1 << 20, 8, 2), "direct, v0 and v1 packed"),
    ("8bit-mixed-case-wild", 8, dict(keyword="Ab*d", wildcard=ord("*")), low_entropy(1 << 20, 6, 3), "direct, packed"),
    ("16bit-kana", 16, dict(keyword="わたし", char_seq=HIRAGANA), low_entropy(1 << 21, 64, 4, bits=16), "direct"),
    ("16bit-mixed-case", 16, dict(keyword="aBc"), low_entropy(1 << 21, 8, 5, bits=16), "hashed"),
    ("16bit-mixed-case-spread", 16, dict(keyword="Abc"), low_entropy(1 << 22, 2048, 6, bits=16), "hashed, many tables"),
    ("8bit-values", 8, dict(values=[1, 2, 3]), low_entropy(1 << 20, 8, 7), "value scan: one (empty) table"),
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_unique_tables_match_the_gui_rule(gpu, case):
    name, bits, pat, data, _ = case
    prog = gpu.Program(bits, keyword=pat.get("keyword"), wildcard=pat.get("wildcard", 0),
                       char_seq=pat.get("char_seq", ()), values=pat.get("values"))
    res = prog.engine_scan(data, 524288)
    _, val = res.arrays()
    assert len(val) > 100, (name, "the case must produce matches to mean anything", len(val))
    expected = gui_unique(prog, val)
    got = res.unique_indices().tolist()
    assert got == expected, (name, len(val), got[:8], expected[:8])
    if pat.get("values") is not None:
        assert got == [0]


def test_unique_of_an_empty_list(gpu):
    prog = gpu.Program(8, keyword="monkey")
    res = prog.engine_scan(np.zeros(4096, np.uint8) + np.arange(4096, dtype=np.uint8) * 0, 524288)
    if res.count == 0:
        assert res.unique_indices().tolist() == []
