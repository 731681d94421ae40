#!/bin/bash
# final validation of round 2 on one B200: GPU test suite, smoke, bench line, ncu captures, bench_search
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > $O/c32_gpu_tests.log 2>&1; echo "pytest rc=$?" >> $O/c32_gpu_tests.log
tail -3 $O/c32_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/c32_smoke.log 2>&1; tail -2 $O/c32_smoke.log
timeout 1500 python bench.py --steps 20 --warmup 3 > $O/c32_bench_n1.json 2> $O/c32_bench_n1.err; echo "rc=$?" >> $O/c32_bench_n1.err
tail -c 300 $O/c32_bench_n1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches_bench_cfg2_final.csv python bench.py --steps 2 --warmup 3 --per-config none --chain off --no-cpu-baseline --no-e2e --no-verify > /dev/null 2>&1
NCU_SKIP=2 bash scripts/ncu_capture.sh k_filter mokeys r2_ncu_k_filter_cfg2_final
NCU_SKIP=2 bash scripts/ncu_capture.sh k_resolve monkey r2_ncu_k_resolve_monkey_final
timeout 600 ./benchmarks/bench_search 0.3 engine > $O/r2_bench_search_gpu_final.txt 2>&1
tail -8 $O/r2_bench_search_gpu_final.txt
python - <<'PY'
import json
d = json.loads(open("gpurun_out/c32_bench_n1.json").read().strip().splitlines()[-1])
print("headline", d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"].get("whole_scan_frac"), d["e2e"]["value"], d["parity"]["bit_exact"])
for p in d.get("per_config", []):
    print(p.get("workload"), p.get("value"), p.get("ms_per_step"), (p.get("roofline") or {}).get("frac"), (p.get("e2e") or {}).get("value"), (p.get("parity") or {}).get("bit_exact"), p.get("error"))
print("chain", d.get("single_chain"))
PY
