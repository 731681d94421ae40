#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2_perf_probe.txt
echo "# scripts/perf_probe.py on the final round-2 library: device-resident engine scans (block 524288), best of 8; filter = the filter kernel's own events" > $O
for mib in 16 512 2048; do
  echo "== $mib MiB" >> $O
  PROBE_ITERS=8 timeout 300 python scripts/perf_probe.py $mib 2>&1 >> $O
done
cat $O
bash scripts/sanitize.sh | tail -12
