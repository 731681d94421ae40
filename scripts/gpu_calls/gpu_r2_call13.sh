#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 900 python bench.py --steps 20 --warmup 3 > $O/c13_bench_n1.json 2> $O/c13_bench_n1.err; echo "rc=$?" >> $O/c13_bench_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/c13_bench_ref.json 2> $O/c13_bench_ref.err; echo "rc=$?" >> $O/c13_bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/c13_launches_bench.csv python bench.py --steps 2 --warmup 1 --per-config none --no-cpu-baseline --no-verify --no-e2e > $O/c13_ncu_bench.log 2>&1
tail -c 300 $O/c13_bench_n1.err; tail -c 300 $O/c13_bench_ref.err
