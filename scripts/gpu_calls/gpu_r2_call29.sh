#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden_fixtures.py -x -q -m gpu > gpurun_out/c29_parity.log 2>&1
echo "parity rc=$?" >> gpurun_out/c29_parity.log
tail -5 gpurun_out/c29_parity.log
O=gpurun_out/c29_l2hint.txt; : > $O
for h in 0 1; do
  echo "== MMG_L2_HINT=$h, 512 MiB" >> $O
  MMG_L2_HINT=$h PROBE_ITERS=8 timeout 300 python scripts/perf_probe.py 512 2>&1 | grep -v distinct >> $O
done
cat $O
D=gpurun_out/c29_dram_writes.txt; : > $D
for h in 0 1; do
  echo "== k_filter 512 MiB, coalesced extent stores, MMG_L2_HINT=$h" >> $D
  MMG_L2_HINT=$h PROBE_ITERS=3 PROBE_CASES="16le abcde" MMG_NO_SPARSE_RESOLVE=1 timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_filter -s 1 -c 1 python scripts/perf_probe.py 512 2>&1 | grep -E "k_filter|dram__" >> $D
done
cat $D
