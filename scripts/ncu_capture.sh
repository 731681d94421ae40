#!/bin/bash
# GPU box: one `ncu --set full` capture of a kernel of scripts/ncu_probe.py, summarised to text (the .ncu-rep stays on the box)
#   scripts/ncu_capture.sh <kernel regex> <probe name> <out name>
cd "$(dirname "$0")/.."
ncu --set full --clock-control none --import-source on -k regex:$1 -s ${NCU_SKIP:-2} -c 1 -o /tmp/$3 -f python scripts/ncu_probe.py $2 ${NCU_SIZE:-512} > /tmp/$3.log 2>&1
python scripts/ncu_summary.py /tmp/$3.ncu-rep ${NCU_MINEX:-0.5} > gpurun_out/$3.txt 2>&1
echo "$3: $(grep gpu__time_duration gpurun_out/$3.txt)"
