#!/bin/bash
# last validation of round 2: whole GPU suite (incl. long keywords, complete-match mode), smoke, compute-sanitizer
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 600 python -m pytest tests/test_long_keywords.py -x -q > gpurun_out/c43_long.log 2>&1; echo "long rc=$?" >> gpurun_out/c43_long.log; tail -12 gpurun_out/c43_long.log
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/c43_gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c43_gpu_tests.log
tail -4 gpurun_out/c43_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
bash scripts/sanitize.sh | tail -12
