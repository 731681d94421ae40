import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import monkey_moore_b200 as mm
from _cases import random_data
from _oracle import Oracle
pat = dict(keyword='eaeeadb*ac*c', wildcard=42)
rng = np.random.default_rng(5)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000003
data = random_data(rng, 8, n, pat)
o = Oracle(8, keyword=pat["keyword"], wildcard=42)
prog = mm.Program(8, keyword=pat["keyword"], wildcard=42)
for rep in range(2):
    r = prog.search(data)
    off = r.offsets
    exp, _ = o.search(data)
    print("rep", rep, "n", n, "got", len(off), "exp", len(exp), "OK" if off.tolist() == exp.tolist() else "FAIL", r.stats(), flush=True)
    if off.tolist() != exp.tolist():
        g, e = set(off.tolist()), set(exp.tolist())
        extra = sorted(g - e)[:10]; missing = sorted(e - g)[:10]
        print(" extra", extra, "missing", missing)
