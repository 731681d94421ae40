#!/bin/bash
# Development aid (GPU box): perf probe for the in-tree library and every variant library
cd "$(dirname "$0")/.."
export PROBE_ITERS=${PROBE_ITERS:-6}
for lib in monkey-moore_b200/libmmoore_b200.so gpurun_variants/lib_*.so; do
  [ -f $lib ] || continue
  echo "== $lib"
  MMG_LIB=$PWD/$lib python scripts/perf_probe.py ${SIZE:-512} 2>&1 | grep -v "^$"
done
