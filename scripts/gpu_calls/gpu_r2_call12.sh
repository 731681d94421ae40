#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/c12_gpu_tests.log 2>&1; echo "pytest rc=$?" >> $O/c12_gpu_tests.log
PROBE_ITERS=8 timeout 300 python scripts/perf_probe.py 16 > $O/c12_probe16.txt 2>&1
MMOORE_PROFILE=1 timeout 300 benchmarks/bench_search 0.2 engine > $O/c12_bench_search_gpu.txt 2> $O/c12_bench_search_gpu.err
tail -3 $O/c12_gpu_tests.log; cat $O/c12_probe16.txt; grep -E "16777216|Engine" $O/c12_bench_search_gpu.txt; tail -4 $O/c12_bench_search_gpu.err
