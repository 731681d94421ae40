"""bench_search-shaped timing: MonkeyMoore::search on one buffer (single chain), device resident and from host."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import monkey_moore_b200 as m
rng = np.random.default_rng(42)
for bits, kw in ((8, dict(keyword="abcde")), (16, dict(keyword="abcde")), (8, dict(keyword="ab*de", wildcard=42)), (8, dict(keyword="monkey"))):
    prog = m.Program(bits, **kw)
    for size in (128 << 10, 2 << 20, 16 << 20, 256 << 20):
        host = rng.integers(0, 256, size, dtype=np.uint8)
        dev = torch.from_numpy(host).cuda()
        harr = host if bits == 8 else host.view(np.uint16)
        darr = dev if bits == 8 else dev.view(torch.int16)
        for _ in range(3):
            prog.search(darr).close()
        torch.cuda.synchronize()
        N = 20
        t0 = time.perf_counter()
        for _ in range(N):
            r = prog.search(darr); n = r.count; r.close()
        t1 = time.perf_counter()
        for _ in range(2):
            prog.search(harr).close()
        t2 = time.perf_counter()
        for _ in range(5):
            r = prog.search(harr); r.close()
        t3 = time.perf_counter()
        print(f"{bits:2d}-bit {kw['keyword']:8s} {size>>10:7d} KiB  device {1e6*(t1-t0)/N:8.1f} us ({size/((t1-t0)/N)/1e9:7.1f} GB/s)   host {1e6*(t3-t2)/5:9.1f} us ({size/((t3-t2)/5)/1e9:6.1f} GB/s) matches={n}", flush=True)
