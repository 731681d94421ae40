#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/c18_gpu_tests.log 2>&1; echo "pytest rc=$?" >> $O/c18_gpu_tests.log
bash scripts/sanitize.sh > /dev/null 2>&1
timeout 900 python bench.py --steps 20 --warmup 3 > $O/c18_bench_n1.json 2> $O/c18_bench_n1.err; echo "rc=$?" >> $O/c18_bench_n1.err
NCU_SKIP=2 bash scripts/ncu_capture.sh k_filter mokeys r2_ncu_k_filter_cfg2_fused
NCU_SKIP=2 bash scripts/ncu_capture.sh k_filter8 monkey r2_ncu_k_filter8_monkey
NCU_SKIP=2 bash scripts/ncu_capture.sh k_filter8 values10 r2_ncu_k_filter8_values10
NCU_SKIP=2 bash scripts/ncu_capture.sh k_resolve monkey r2_ncu_k_resolve_monkey
tail -3 $O/c18_gpu_tests.log; cat $O/r2_sanitizer.txt; tail -c 200 $O/c18_bench_n1.err
