#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/c7_gpu_tests.log 2>&1; echo "pytest rc=$?" >> $O/c7_gpu_tests.log
timeout 300 python scripts/perf_probe.py 512 > $O/c7_probe.txt 2>&1
tail -3 $O/c7_gpu_tests.log; cat $O/c7_probe.txt
