#!/bin/bash
mkdir -p gpurun_out
MMG_LIB=$PWD/monkey-moore_b200/libmmoore_b200_prof.so timeout 300 python scripts/resolve_phases.py 16 > gpurun_out/c23_phases16.txt 2>&1
cat gpurun_out/c23_phases16.txt
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden_fixtures.py tests/test_gpu_unique.py -x -q -m gpu > gpurun_out/c23_parity.log 2>&1
echo "parity rc=$?" >> gpurun_out/c23_parity.log
tail -5 gpurun_out/c23_parity.log
PROBE_ITERS=8 timeout 300 python scripts/perf_probe.py 16 > gpurun_out/c23_probe16.txt 2>&1
cat gpurun_out/c23_probe16.txt
PROBE_ITERS=6 PROBE_CASES="8 " timeout 300 python scripts/perf_probe.py 512 > gpurun_out/c23_probe512.txt 2>&1
cat gpurun_out/c23_probe512.txt
