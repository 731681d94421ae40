#!/bin/bash
cd "$(dirname "$0")/.."
export PROBE_ITERS=6 PROBE_CASES=${PROBE_CASES:-16}
for lib in monkey-moore_b200/libmmoore_b200.so gpurun_variants/lib_*.so; do
  [ -f $lib ] || continue
  for cpw in 8 4 2; do
    echo "== $lib chunks_per_warp>=$cpw"
    MMG_LIB=$PWD/$lib MMG_CHUNKS_PER_WARP=$cpw python scripts/perf_probe.py 512 2>&1 | grep -v "^$"
  done
done
