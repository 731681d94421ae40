#!/bin/bash
# chain slices (single GPU) + static chunk scheduling check at 16 MiB
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "chain" > gpurun_out/c21_chain.log 2>&1
echo "chain rc=$?" >> gpurun_out/c21_chain.log
tail -30 gpurun_out/c21_chain.log
PROBE_ITERS=8 timeout 300 python scripts/perf_probe.py 16 > gpurun_out/c21_probe16.txt 2>&1
cat gpurun_out/c21_probe16.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "not chain" > gpurun_out/c21_parity.log 2>&1
echo "parity rc=$?" >> gpurun_out/c21_parity.log
tail -5 gpurun_out/c21_parity.log
