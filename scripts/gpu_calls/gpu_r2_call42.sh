#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 600 python -m pytest tests/test_complete_mode.py -x -q > gpurun_out/c42_complete.log 2>&1
echo "complete rc=$?" >> gpurun_out/c42_complete.log; tail -15 gpurun_out/c42_complete.log
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_golden_fixtures.py -x -q -m gpu > gpurun_out/c42_parity.log 2>&1
echo "parity rc=$?" >> gpurun_out/c42_parity.log; tail -3 gpurun_out/c42_parity.log
