#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/c30_stage_sweep.txt; : > $O
nproc >> $O
for mib in 32 64 256 1024; do
  for mode in 4 100000; do
    echo "== $mib MiB, MMG_STAGE_MIN_MIB=$mode ($( [ $mode = 4 ] && echo ring || echo direct ))" >> $O
    MMG_PROFILE_COPY=1 MMG_STAGE_MIN_MIB=$mode timeout 120 python scripts/search_host_probe.py $mib 2>&1 | tail -4 >> $O
  done
done
cat $O
