"""Development aid: per-phase time of k_resolve (library built with -DMMG_RESOLVE_PROF), summed over CTAs."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import monkey_moore_b200 as m
size = (int(sys.argv[1]) if len(sys.argv) > 1 else 512) << 20
torch.manual_seed(0)
low = torch.randint(0, 16, (size,), dtype=torch.uint8, device="cuda")
data = torch.randint(0, 256, (size,), dtype=torch.uint8, device="cuda")
lib = m.lib()
out = (ctypes.c_ulonglong * 8)()
for name, buf, pat in [("abc low16", low, dict(keyword="abc")), ("values10", data, dict(values=[10, 12, 15, 11, 30, 31, 29, 40, 41, 45])),
                       ("monkey", data, dict(keyword="monkey"))]:
    prog = m.Program(8, **pat)
    for it in range(3):
        lib.mmg_debug_resolve_phases(out, 1)
        r = prog.engine_scan(buf, 524288); st = r.stats(); n = r.count; r.close()
        lib.mmg_debug_resolve_phases(out, 0)
    nb = size // 524288
    print(name, "matches", n, "events", st["events"], "total %.3f filter %.3f ms" % (st["ms_total"], st["ms_filter"]),
          "per-CTA avg us by phase [setup, a, b, c, d-wait?, e...]:", ["%.1f" % (out[i] / nb / 1e3) for i in range(7)])
