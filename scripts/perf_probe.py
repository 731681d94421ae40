"""Quick device-resident timing of the scan on a few patterns (development aid, not the bench)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import monkey_moore_b200 as m

MiB = 1 << 20
size = int(sys.argv[1]) * MiB if len(sys.argv) > 1 else 512 * MiB
torch.manual_seed(0)
data = torch.randint(0, 256, (size,), dtype=torch.uint8, device="cuda")
low = torch.randint(0, 16, (size,), dtype=torch.uint8, device="cuda")
torch.cuda.synchronize()
cases = [
    ("16le mo*key*s", 16, dict(keyword="mo*key*s", wildcard=ord("*")), data, False),
    ("16be mo*key*s", 16, dict(keyword="mo*key*s", wildcard=ord("*")), data, True),
    ("16le abcde", 16, dict(keyword="abcde"), data, False),
    ("16le kana6", 16, dict(keyword="わたしたちは", char_seq="あいうえおかきくけこさしすせそたちつてとなにぬねのはひふへほまみむめもやゆよらりるれろわをゃっゅょ"), data, False),
    ("8 monkey", 8, dict(keyword="monkey"), data, False),
    ("8 abcde", 8, dict(keyword="abcde"), data, False),
    ("8 ab*de", 8, dict(keyword="ab*de", wildcard=ord("*")), data, False),
    ("8 values10", 8, dict(values=[10, 12, 15, 11, 30, 31, 29, 40, 41, 45]), data, False),
    ("8 abc low16", 8, dict(keyword="abc"), low, False),
]
sel = os.environ.get("PROBE_CASES")
if sel:
    cases = [c for c in cases if any(c[0].startswith(t) for t in sel.split(","))]
for name, bits, pat, buf, be in cases:
    prog = m.Program(bits, **pat)
    best = None
    for it in range(int(os.environ.get('PROBE_ITERS', '4'))):
        r = prog.engine_scan(buf, 524288, big_endian=be)
        st = r.stats(); n = r.count; r.close()
        if best is None or st["ms_total"] < best["ms_total"]:
            best = st
    gbs = size / best["ms_total"] / 1e6
    gbf = size / best["ms_filter"] / 1e6
    print(f"{name:16s} matches={n:9d} events={best['events']:10d} total={best['ms_total']:8.3f} ms ({gbs:7.1f} GB/s) "
          f"filter={best['ms_filter']:8.3f} ms ({gbf:7.1f} GB/s) launches={best['launches']}", flush=True)
    if name.startswith("8 abc low16"):
        import time
        r = prog.engine_scan(buf, 524288)
        r.count
        torch.cuda.synchronize(); t0 = time.perf_counter()
        u = r.unique_indices()
        torch.cuda.synchronize(); t1 = time.perf_counter()
        print(f"   distinct tables: {len(u)} of {r.count} matches in {(t1 - t0) * 1e3:.3f} ms (mmg_results_unique, host call)")
        r.close()
