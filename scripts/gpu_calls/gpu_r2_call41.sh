#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden_fixtures.py tests/test_gpu_unique.py tests/test_cpp_layer.py -x -q -m gpu > gpurun_out/c41_parity.log 2>&1
echo "parity rc=$?" >> gpurun_out/c41_parity.log; tail -4 gpurun_out/c41_parity.log
PROBE_ITERS=8 PROBE_CASES="8 " timeout 300 python scripts/perf_probe.py 512 2>&1 | grep -v distinct
PROBE_ITERS=4 PROBE_CASES="8 abc,8 ab" timeout 300 python scripts/perf_probe.py 2048 2>&1 | grep -v distinct
