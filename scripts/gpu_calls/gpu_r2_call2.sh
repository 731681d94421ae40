#!/bin/bash
# round 2, call 2 (1 GPU): parity suite with the slotted 8-bit filter + device-resident probe (slotted vs list format)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/c2_gpu_tests.log 2>&1; echo "pytest rc=$?" >> $O/c2_gpu_tests.log
PROBE_CASES="8 " timeout 300 python scripts/perf_probe.py 512 > $O/c2_probe_slotted.txt 2>&1
PROBE_CASES="8 " MMG_NO_SLOTTED=1 timeout 300 python scripts/perf_probe.py 512 > $O/c2_probe_list.txt 2>&1
tail -5 $O/c2_gpu_tests.log; cat $O/c2_probe_slotted.txt; cat $O/c2_probe_list.txt
