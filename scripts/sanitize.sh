#!/bin/bash
# GPU box: compute-sanitizer over scripts/sanitize_target.py -> gpurun_out/r2_sanitizer.txt
cd "$(dirname "$0")/.."
O=gpurun_out/r2_sanitizer.txt
echo "# compute-sanitizer over scripts/sanitize_target.py, library built from $(sha256sum monkey-moore_b200/csrc/scan_kernels.cu | cut -c1-16) (scan_kernels.cu) $(date -u +%FT%TZ)" > $O
for tool in memcheck racecheck synccheck; do
    echo "== $tool" >> $O
    timeout 900 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_target.py > /tmp/san_$tool.log 2>&1
    echo "exit code $?" >> $O
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize_target:|Error|error|Invalid|hazard" /tmp/san_$tool.log | head -30 >> $O
done
cat $O
