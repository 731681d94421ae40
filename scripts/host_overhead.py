"""Host-side cost of one scan: enqueue time of the async call, completion time, and back-to-back throughput."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import monkey_moore_b200 as m

for mib in (16, 512):
    size = mib << 20
    data = torch.randint(0, 256, (size,), dtype=torch.uint8, device="cuda")
    prog = m.Program(16, keyword="mo*key*s", wildcard=ord("*"))
    for _ in range(5):
        prog.engine_scan(data, 524288).close()
    torch.cuda.synchronize()
    N = 200
    t_enq = t_fin = 0.0
    t0 = time.perf_counter()
    for _ in range(N):
        a = time.perf_counter()
        r = prog.engine_scan(data, 524288, asynchronous=True)
        b = time.perf_counter()
        r.stats()
        c = time.perf_counter()
        r.close()
        t_enq += b - a; t_fin += c - b
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    # pipelined: keep 4 scans in flight
    q = []
    t2 = time.perf_counter()
    for _ in range(N):
        q.append(prog.engine_scan(data, 524288, asynchronous=True))
        if len(q) >= 4:
            q.pop(0).close()
    for r in q:
        r.close()
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    print(f"{mib:4d} MiB: enqueue {t_enq/N*1e6:6.1f} us  wait+stats {t_fin/N*1e6:6.1f} us  serial {1e6*(t1-t0)/N:6.1f} us/scan "
          f"({size/((t1-t0)/N)/1e9:7.1f} GB/s)  pipelined x4 {1e6*(t3-t2)/N:6.1f} us/scan ({size/((t3-t2)/N)/1e9:7.1f} GB/s)")
