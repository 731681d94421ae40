#!/bin/bash
# final validation of round 2 on one B200 (after the split resolve): GPU test suite, smoke, sanitizer, bench line
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > $O/c39_gpu_tests.log 2>&1; echo "pytest rc=$?" >> $O/c39_gpu_tests.log
tail -3 $O/c39_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/c39_smoke.log 2>&1; tail -2 $O/c39_smoke.log
bash scripts/sanitize.sh | tail -12
timeout 1500 python bench.py --steps 20 --warmup 3 > $O/c39_bench_n1.json 2> $O/c39_bench_n1.err; echo "rc=$?" >> $O/c39_bench_n1.err
tail -c 300 $O/c39_bench_n1.err
PROBE_ITERS=8 timeout 300 python scripts/perf_probe.py 16 > $O/c39_probe16.txt 2>&1
python - <<'PY'
import json
d = json.loads(open("gpurun_out/c39_bench_n1.json").read().strip().splitlines()[-1])
print("headline", d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"].get("whole_scan_frac"), d["roofline"]["traffic"], d["e2e"]["value"], d["parity"]["bit_exact"])
for p in d.get("per_config", []):
    print(p.get("workload"), p.get("value"), p.get("ms_per_step"), (p.get("roofline") or {}).get("frac"), (p.get("e2e") or {}).get("value"), (p.get("parity") or {}).get("bit_exact"), p.get("error"))
print("chain", d.get("single_chain"))
PY
