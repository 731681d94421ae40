"""Short single-workload driver for ncu captures: python scripts/ncu_target.py cfg2 [size_mib] [search_index]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import monkey_moore_b200 as mm
import monkey_moore_b200.workloads as wl

key = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
w = wl.WORKLOADS[key]
if len(sys.argv) > 2:
    w = w.scaled(int(sys.argv[2]) << 20)
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
blob = wl.device_blob(w)
s = w.searches[which]
prog = mm.Program(w.bits, **s.pattern)
for _ in range(3):
    r = prog.engine_scan(blob, w.block_size, big_endian=s.big_endian)
    print(s.name, r.count, r.stats())
    r.close()
