#!/bin/bash
# diagnostics: staging threshold sweep, DRAM write bytes of k_filter vs input size / cache control, compute-sanitizer
mkdir -p gpurun_out
O=gpurun_out/c27_stage_sweep.txt; : > $O
for mib in 8 16 32 64 128 256; do
  for mode in 4 100000; do
    echo "== $mib MiB, MMG_STAGE_MIN_MIB=$mode ($( [ $mode = 4 ] && echo ring || echo direct ))" >> $O
    MMG_STAGE_MIN_MIB=$mode timeout 120 python scripts/search_host_probe.py $mib 2>&1 | tail -3 >> $O
  done
done
cat $O
D=gpurun_out/c27_dram_writes.txt; : > $D
for mib in 512 2048; do
  for cc in all none; do
    echo "== $mib MiB, --cache-control $cc" >> $D
    PROBE_ITERS=3 PROBE_CASES="16le abcde" MMG_NO_SPARSE_RESOLVE=1 timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes_equiv_l1sectormiss_pipe_lsu_mem_global_op_st.sum --cache-control $cc --clock-control none -k regex:k_filter -s 1 -c 2 python scripts/perf_probe.py $mib 2>&1 | grep -E "k_filter|dram__|lts__" >> $D
  done
done
cat $D
bash scripts/sanitize.sh
