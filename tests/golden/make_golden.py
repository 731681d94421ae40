"""Generates the golden fixtures of tests/golden/: outputs of the UNMODIFIED reference (oracle/_ref/libmmref.so, built by
oracle/Makefile from /root/reference/src/core where that checkout exists) on small instances of the BASELINE
configurations.  Run where /root/reference is present:

    python tests/golden/make_golden.py

For every workload of monkey_moore_b200.workloads (cfg1..cfg5, scaled to SIZE bytes) and every search of its step the
reference's mmoore::SearchEngine<T>::run (block 524288, 4 threads) scans the synthetic blob from a file; the fixture
stores the match offsets, the inferred table of every match (sorted keys + one row of values per match) and a SHA-256
of the blob, so the tests also pin the synthetic generator.  tests/test_golden_fixtures.py compares the C restatement
(CPU) and the CUDA path (GPU) with these files; nothing there needs /root/reference or libmmref.so.
"""
import hashlib
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

SIZE = {"cfg1": 2 << 20, "cfg2": 2 << 20, "cfg3": 2 << 20, "cfg4": 2 << 20, "cfg5": 1 << 20}


def main():
    import monkey_moore_b200.workloads as wl
    from _oracle import Ref
    assert Ref.available(), "oracle/_ref/libmmref.so is missing: build it where /root/reference exists (make -C oracle)"
    tmp = tempfile.mkdtemp()
    for key, size in SIZE.items():
        w = wl.WORKLOADS[key].scaled(size)
        blob = wl.host_blob(w)
        path = os.path.join(tmp, key + ".bin")
        blob.tofile(path)
        digest = hashlib.sha256(blob.tobytes()).hexdigest()
        for s in w.searches:
            p = s.pattern
            r = Ref.engine(w.bits, path, keyword=p.get("keyword"), wildcard=p.get("wildcard", 0),
                           char_seq=p.get("char_seq", ()), values=p.get("values"), big_endian=s.big_endian,
                           threads=4, block=w.block_size)
            maps = r["maps"]
            keys = sorted(maps[0]) if maps else []
            assert all(sorted(m) == keys for m in maps)
            table = np.array([[m[k] for k in keys] for m in maps], dtype=np.uint32).reshape(len(maps), len(keys))
            out = os.path.join(HERE, "%s_%s.npz" % (key, s.name.replace("*", "x")))
            np.savez_compressed(out, offsets=r["offsets"].astype(np.uint64), keys=np.array(keys, dtype=np.uint32),
                                table=table, blob_sha256=np.array(digest), size=np.array(size, dtype=np.uint64),
                                big_endian=np.array(s.big_endian))
            print("%-32s matches=%6d table=%2d entries  %s" % (os.path.basename(out), len(maps), len(keys), digest[:12]))


if __name__ == "__main__":
    main()
