"""Bounded workload for compute-sanitizer (memcheck / racecheck / synccheck): every kernel of the shipped library on small
inputs -- the reference's known-answer vectors on all code paths, the five BASELINE configurations at 1-2 MiB (list
format, fused sparse resolve and its hand-over), the segmented path of large blocks, search() on one chain -- whole
and in slices (mmg_chain_*) --, the distinct-table reduction.  Results are checked against the oracle, so a sanitizer run is also a parity run.
    compute-sanitizer --tool memcheck python scripts/sanitize_target.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

import monkey_moore_b200 as mm
import monkey_moore_b200.workloads as wl
from _cases import ref_kats
from _oracle import Oracle


def kw(p):
    return dict(keyword=p.get("keyword"), wildcard=p.get("wildcard", 0), char_seq=p.get("char_seq", ()), values=p.get("values"))


n = 0
for kat in ref_kats():
    for override in (0, 1, 2):
        old = mm.set_path_override(override)
        try:
            prog = mm.Program(kat["bits"], **kw(kat))
            r = prog.search(kat["data"])
            assert r.offsets.tolist() == list(kat["pos"]), kat["name"]
            if kat["maps"] is not None:
                assert r.tables() == kat["maps"], kat["name"]
            r.close()
            n += 1
        finally:
            mm.set_path_override(old)
for key, size in (("cfg1", 1 << 20), ("cfg2", 2 << 20), ("cfg3", 1 << 20), ("cfg4", 2 << 20), ("cfg5", 1 << 20)):
    w = wl.WORKLOADS[key].scaled(size)
    blob = wl.host_blob(w)
    for s in w.searches:
        prog = mm.Program(w.bits, **s.pattern)
        o = Oracle(w.bits, **kw(s.pattern))
        for block in (w.block_size, 65536, 1 << 20):          # 1 MiB blocks: more than 128 sub-tiles -> segmented resolve
            exp, expv = o.engine(blob, block, big_endian=s.big_endian, wrap32=False)
            for rep in range(3):                                # scans 2 and 3 of a sparse pattern take the fused resolve
                r = prog.engine_scan(blob, block, big_endian=s.big_endian)
                off, val = r.arrays()
                assert off.tolist() == exp.tolist() and val.tolist() == expv.tolist(), (key, s.name, block, rep)
                if rep == 2 and len(off):
                    r.unique_indices()
                r.close()
                n += 1
        data = blob if w.bits == 8 else blob[: len(blob) // 2 * 2].view(np.uint16)
        r = prog.search(data)
        pos, _ = o.search(data if not s.big_endian else data)
        if not s.big_endian:
            assert r.offsets.tolist() == pos.tolist(), (key, s.name, "search")
        r.close()
        n += 1
        if not s.big_endian:
            # the same chain in slices (mmg_chain_*): one sub-tile, 64 and 128 sub-tiles (the latter adds the overlap segment)
            W = w.bits // 8
            for subs, view in ((1, data[: (5 * 4096 + 78) // W]), (64, data), (128, data)):
                want = pos if view is data else o.search(view)[0]
                parts = prog.search_sliced(view, subs * 4096 // W)
                got = np.concatenate([p.arrays()[0] for p in parts])
                for p in parts:
                    p.close()
                assert got.tolist() == want.tolist(), (key, s.name, "sliced", subs)
                n += 1
print("sanitize_target: %d scans, all bit-exact" % n)
