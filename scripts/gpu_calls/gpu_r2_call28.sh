#!/bin/bash
# L2 hint A/B for the input stream + does a pure read kernel show DRAM writes too?
mkdir -p gpurun_out
O=gpurun_out/c28_l2hint.txt; : > $O
for h in 0 1 2; do
  echo "== MMG_L2_HINT=$h, 512 MiB" >> $O
  MMG_L2_HINT=$h PROBE_ITERS=8 timeout 300 python scripts/perf_probe.py 512 2>&1 | grep -v distinct >> $O
done
for h in 0 1; do
  echo "== MMG_L2_HINT=$h, 16 MiB" >> $O
  MMG_L2_HINT=$h PROBE_ITERS=8 PROBE_CASES="8 monkey,16le mo,8 abc" timeout 300 python scripts/perf_probe.py 16 2>&1 | grep -v distinct >> $O
  echo "== MMG_L2_HINT=$h, 2048 MiB" >> $O
  MMG_L2_HINT=$h PROBE_ITERS=4 PROBE_CASES="8 monkey,16le mo,8 abc" timeout 300 python scripts/perf_probe.py 2048 2>&1 | grep -v distinct >> $O
done
cat $O
D=gpurun_out/c28_dram_writes.txt; : > $D
echo "== pure read kernel (scripts/micro/readbw, k_ldg), 512 MiB" >> $D
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_ldg -c 2 scripts/micro/readbw 512 2>&1 | grep -E "k_ldg|dram__" >> $D
for h in 0 1; do
  echo "== k_filter 512 MiB, MMG_L2_HINT=$h" >> $D
  MMG_L2_HINT=$h PROBE_ITERS=3 PROBE_CASES="16le abcde" MMG_NO_SPARSE_RESOLVE=1 timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_filter -s 1 -c 1 python scripts/perf_probe.py 512 2>&1 | grep -E "k_filter|dram__" >> $D
done
cat $D
