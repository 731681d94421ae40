#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
NCU_SIZE=16 NCU_SKIP=2 NCU_MINEX=0.002 bash scripts/ncu_capture.sh k_resolve monkey c38_ncu_resolve_16m_split
grep -E "gpu__time|warps_active|issue_active|stalled_(barrier|long|short|wait|membar|lg|no_inst)|inst_executed.sum|launch__" gpurun_out/c38_ncu_resolve_16m_split.txt
sed -n '/hot (/,$p' gpurun_out/c38_ncu_resolve_16m_split.txt | head -70
