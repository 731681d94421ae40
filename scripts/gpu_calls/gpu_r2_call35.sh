#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/c35_chunks_sweep.txt; : > $O
for cpw in 2 4 8 16; do
  echo "== MMG_CHUNKS_PER_WARP=$cpw, 512 MiB" >> $O
  MMG_CHUNKS_PER_WARP=$cpw PROBE_ITERS=8 PROBE_CASES="16le mo,16le abcde,16le kana,8 monkey,8 abcde,8 abc" timeout 300 python scripts/perf_probe.py 512 2>&1 | grep -v distinct >> $O
done
cat $O
NCU_SIZE=2048 NCU_SKIP=2 NCU_MINEX=0.5 bash scripts/ncu_capture.sh k_resolve abclow16 c35_ncu_resolve_abclow16_2g
head -80 gpurun_out/c35_ncu_resolve_abclow16_2g.txt
