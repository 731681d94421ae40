"""One named pattern on a random (or low-entropy) device blob, 3 scans: the target of ncu captures.
   python scripts/ncu_probe.py monkey|abcde|values10|abclow16|abclow4|abwde [size_mib]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import monkey_moore_b200 as m

name = sys.argv[1] if len(sys.argv) > 1 else "monkey"
size = (int(sys.argv[2]) if len(sys.argv) > 2 else 512) << 20
torch.manual_seed(0)
hi = {"abclow16": 16, "abclow4": 4}.get(name, 256)
data = torch.randint(0, hi, (size,), dtype=torch.uint8, device="cuda")
pats = {
    "monkey": dict(keyword="monkey"), "abcde": dict(keyword="abcde"), "abwde": dict(keyword="ab*de", wildcard=ord("*")),
    "values10": dict(values=[10, 12, 15, 11, 30, 31, 29, 40, 41, 45]), "abclow16": dict(keyword="abc"), "abclow4": dict(keyword="abc"),
}
KANA = "あいうえおかきくけこさしすせそたちつてとなにぬねのはひふへほまみむめもやゆよらりるれろわをゃっゅょ"
pats16 = {"kana6": dict(keyword="わたしたちは", char_seq=KANA), "mokeys": dict(keyword="mo*key*s", wildcard=ord("*"))}
prog = m.Program(16, **pats16[name]) if name in pats16 else m.Program(8, **pats[name])
for _ in range(3):
    r = prog.engine_scan(data, 524288)
    print(name, r.count, r.stats())
    r.close()
