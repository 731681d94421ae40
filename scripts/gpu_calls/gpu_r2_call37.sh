#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden_fixtures.py tests/test_gpu_unique.py tests/test_cpp_layer.py -x -q -m gpu > gpurun_out/c37_parity.log 2>&1
echo "parity rc=$?" >> gpurun_out/c37_parity.log; tail -4 gpurun_out/c37_parity.log
MMG_LIB=$PWD/monkey-moore_b200/libmmoore_b200_prof.so timeout 300 python scripts/resolve_phases.py 16 > gpurun_out/c37_phases16.txt 2>&1; cat gpurun_out/c37_phases16.txt
for e in 0 1; do
  echo "== MMG_NO_RESOLVE_SPLIT=$e"
  if [ $e = 1 ]; then export MMG_NO_RESOLVE_SPLIT=1; fi
  PROBE_ITERS=8 PROBE_CASES="8 " timeout 300 python scripts/perf_probe.py 16 2>&1 | grep -v distinct
  PROBE_ITERS=8 PROBE_CASES="8 " timeout 300 python scripts/perf_probe.py 128 2>&1 | grep -v distinct
done
