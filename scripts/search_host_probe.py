"""Host-pointer search() timing (pageable numpy buffer): python scripts/search_host_probe.py [MiB]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import monkey_moore_b200 as m
n = (int(sys.argv[1]) if len(sys.argv) > 1 else 16) << 20
data = np.random.default_rng(0).integers(0, 256, n, dtype=np.uint8)
prog = m.Program(8, keyword="monkey")
for it in range(6):
    t = time.perf_counter()
    r = prog.search(data); c = r.count; st = r.stats(); r.close()
    dt = time.perf_counter() - t
    print("search %d MiB pageable: %.3f ms  %.2f GB/s  (h2d %.3f ms, scan %.3f ms)" % (n >> 20, dt * 1e3, n / dt / 1e9, st["ms_h2d"], st["ms_total"]), flush=True)
