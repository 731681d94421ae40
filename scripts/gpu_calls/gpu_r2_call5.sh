#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/c5_gpu_tests.log 2>&1; echo "pytest rc=$?" >> $O/c5_gpu_tests.log
PROBE_CASES="8 " timeout 300 python scripts/perf_probe.py 512 > $O/c5_probe_slotted.txt 2>&1
bash scripts/ncu_capture.sh k_filter8s monkey c5_ncu_filter8s_monkey
bash scripts/ncu_capture.sh k_resolve monkey c5_ncu_resolve_slot_monkey
bash scripts/ncu_capture.sh k_resolve mokeys c5_ncu_resolve_mokeys
tail -3 $O/c5_gpu_tests.log; cat $O/c5_probe_slotted.txt
