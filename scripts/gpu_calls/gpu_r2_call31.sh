#!/bin/bash
# engine ingestion: phases, slab size
mkdir -p gpurun_out
O=gpurun_out/c31_engine.txt; : > $O
nproc >> $O
for slab in 32 16 64; do
  echo "== MMOORE_SLAB_MIB=$slab" >> $O
  MMOORE_SLAB_MIB=$slab MMOORE_PROFILE=1 timeout 300 ./benchmarks/bench_search 0.3 engine-only >> $O 2>&1
done
grep -E "==|SearchEngine|run\(\)" $O | awk '/run\(\)/{c++; if (c%4==0) print; next} {print}' | head -80
