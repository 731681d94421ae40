#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
bash scripts/sanitize.sh | tail -14
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden_fixtures.py -x -q -m gpu > gpurun_out/c40_parity.log 2>&1
echo "parity rc=$?" >> gpurun_out/c40_parity.log; tail -3 gpurun_out/c40_parity.log
