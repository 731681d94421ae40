#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/c6_gpu_tests.log 2>&1; echo "pytest rc=$?" >> $O/c6_gpu_tests.log
PROBE_CASES="8 " timeout 300 python scripts/perf_probe.py 512 > $O/c6_probe_slotted.txt 2>&1
bash scripts/ncu_capture.sh k_filter8s monkey c6_ncu_filter8s_monkey
bash scripts/ncu_capture.sh k_resolve monkey c6_ncu_resolve_slot_monkey
tail -3 $O/c6_gpu_tests.log; cat $O/c6_probe_slotted.txt
