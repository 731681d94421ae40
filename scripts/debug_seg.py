import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import monkey_moore_b200 as mm
from _cases import random_data, random_pattern
from _oracle import Oracle, MMError

def kw(p):
    return dict(keyword=p.get("keyword"), wildcard=p.get("wildcard", 0), char_seq=p.get("char_seq", ()), values=p.get("values"))

for seed in range(4):
    rng = np.random.default_rng(2000 + seed)
    for it in range(12):
        bits = int(rng.choice([8, 16]))
        pat = random_pattern(rng, bits)
        try:
            o = Oracle(bits, **kw(pat))
        except MMError:
            continue
        n = int(rng.choice([70000, 262144, 1000003]))
        data = random_data(rng, bits, n, pat)
        prog = mm.Program(bits, **kw(pat))
        off = prog.search(data).offsets
        exp, _ = o.search(data)
        ok = off.tolist() == exp.tolist()
        W = bits // 8
        msg = ""
        if not ok:
            k = 0
            while k < min(len(off), len(exp)) and off[k] == exp[k]:
                k += 1
            a = int(off[k]) if k < len(off) else -1
            b = int(exp[k]) if k < len(exp) else -1
            msg = " first diff idx %d got %d exp %d (subtile %d / %d, seg %d / %d) len %d vs %d" % (
                k, a, b, a * W // 4096, b * W // 4096, a * W // (4096 * 128), b * W // (4096 * 128), len(off), len(exp))
        print(seed, it, bits, n, {k_: (v if not isinstance(v, (list, tuple)) or len(v) < 12 else "...") for k_, v in pat.items()}, "OK" if ok else "FAIL" + msg, flush=True)
