#!/bin/bash
mkdir -p gpurun_out
MMG_LIB=$PWD/monkey-moore_b200/libmmoore_b200_prof.so timeout 300 python scripts/resolve_phases.py 16 > gpurun_out/c24_phases16.txt 2>&1
cat gpurun_out/c24_phases16.txt
PROBE_ITERS=8 PROBE_CASES="8 " timeout 300 python scripts/perf_probe.py 16 > gpurun_out/c24_probe16.txt 2>&1
cat gpurun_out/c24_probe16.txt
PROBE_ITERS=6 PROBE_CASES="8 " timeout 300 python scripts/perf_probe.py 512 > gpurun_out/c24_probe512.txt 2>&1
cat gpurun_out/c24_probe512.txt
PROBE_ITERS=4 PROBE_CASES="8 abc,8 monkey" timeout 300 python scripts/perf_probe.py 2048 > gpurun_out/c24_probe2048.txt 2>&1
cat gpurun_out/c24_probe2048.txt
