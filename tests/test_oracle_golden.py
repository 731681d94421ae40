"""The oracle (oracle/mm_oracle.c) against every known-answer vector the reference's tests hold for this path."""
import numpy as np
import pytest

from _cases import (ENGINE8_BLOCKS, ENGINE8_FILE, ENGINE8_OFFSETS, ENGINE16_BLOCKS_BE, ENGINE16_BLOCKS_LE,
                    ENGINE16_FILE, ENGINE16_OFFSETS, ref_kats)
from _oracle import MMError, Oracle


def kw(k):
    return dict(keyword=k.get("keyword"), wildcard=k.get("wildcard", 0), char_seq=k.get("char_seq", ()),
                values=k.get("values"))


@pytest.mark.parametrize("kat", ref_kats(), ids=lambda k: k["name"])
def test_search_known_answers(kat):
    """/root/reference/tests/test_monkey_moore.cpp:13-344"""
    o = Oracle(kat["bits"], **kw(kat))
    pos, vals = o.search(kat["data"])
    assert pos.tolist() == kat["pos"]
    if kat["maps"] is not None:
        assert [o.table(int(v[0]), int(v[1])) for v in vals] == kat["maps"]


def test_engine_known_answers():
    """/root/reference/tests/test_search_engine.cpp:26-158 (threads x block sizes, LE and BE)"""
    for bits, file, offs, blocks, be in [(8, ENGINE8_FILE, ENGINE8_OFFSETS, ENGINE8_BLOCKS, False),
                                         (16, ENGINE16_FILE, ENGINE16_OFFSETS, ENGINE16_BLOCKS_LE, False),
                                         (16, ENGINE16_FILE.byteswap(), ENGINE16_OFFSETS, ENGINE16_BLOCKS_BE, True)]:
        fb = np.ascontiguousarray(file).view(np.uint8)
        o = Oracle(bits, keyword="text", wildcard=ord("*"))
        for b in blocks:
            for wrap in (True, False):
                assert o.engine(fb, b, big_endian=be, wrap32=wrap)[0].tolist() == offs


def test_engine_wildcard_passthrough_and_counts():
    """tests/test_search_engine.cpp:429-447: '$atch' with wildcard '$' -> 7 results"""
    text = "match#catch#batch#match#patch#hatch#match"
    fb = np.array([(ord(c) - 0x15) & 0xFF for c in text], dtype=np.uint8)
    o = Oracle(8, keyword="$atch", wildcard=ord("$"))
    assert len(o.engine(fb, 20)[0]) == 7


def test_lossy_chain_semantics():
    """constant data with 'aaa' reports 0, 2, 4, ... (not every matching window)."""
    o = Oracle(8, keyword="aaa")
    pos, _ = o.search(np.full(11, 7, np.uint8))
    assert pos.tolist() == [0, 2, 4, 6, 8]


def test_rejections():
    with pytest.raises(MMError):
        Oracle(8, keyword="a")            # advance 0: the reference never terminates
    with pytest.raises(MMError):
        Oracle(8, keyword="***", wildcard=ord("*"))
    with pytest.raises(MMError):
        Oracle(8, keyword=[10, 1000])      # difference does not fit the 8-bit skip table
    Oracle(16, keyword=[10, 1000])
