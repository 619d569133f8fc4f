#!/bin/bash
# usage (on the GPU box): tools/ncu_capture.sh <tag> [model] [batch]
# One `ncu --set full` capture of every kernel of a forward (the SECOND forward of tools/one_forward.py), exported
# as the raw CSV page to gpurun_out/<tag>_raw.csv; the .ncu-rep stays on the box (too large to bring back).
TAG=$1; MODEL=${2:-AASIST}; B=${3:-512}; T=/tmp/$(basename $TAG)
set -e
timeout 120 python tools/one_forward.py $MODEL $B > gpurun_out/${TAG}_plain.log 2>&1      # must exit 0 without ncu first
N=$(grep -o "launches [0-9]*" gpurun_out/${TAG}_plain.log | awk '{print $2}')
PER=$((N / 2))
timeout 900 ncu --set full --clock-control none --import-source on -f -o $T python tools/one_forward.py $MODEL $B > gpurun_out/${TAG}_ncu.log 2>&1
ncu -i $T.ncu-rep --page raw --csv > ${T}_all.csv
# header (2 lines) + the last PER kernels = the second forward
(head -2 ${T}_all.csv; tail -n $PER ${T}_all.csv) > gpurun_out/${TAG}_raw.csv
echo "kernels per forward: $PER"; wc -l gpurun_out/${TAG}_raw.csv
