#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "chain" > gpurun_out/c22_chain.log 2>&1
echo "chain rc=$?" >> gpurun_out/c22_chain.log
tail -30 gpurun_out/c22_chain.log
MMG_LIB=$PWD/monkey-moore_b200/libmmoore_b200_prof.so timeout 300 python scripts/resolve_phases.py 16 > gpurun_out/c22_phases16.txt 2>&1
cat gpurun_out/c22_phases16.txt
